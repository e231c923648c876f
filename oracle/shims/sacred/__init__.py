"""Minimal stand-in for the `sacred` package — TEST INFRASTRUCTURE ONLY.

The reference (Jarvis73/PEMP) wires its hyper-parameters through Sacred
ingredients (`networks/pemp_stage1.py:12-37`, `networks/baseline.py:11-33`,
`networks/panet.py:11-32`).  Sacred is not installed in this image, so the
oracle imports the reference through this shim.  It reproduces only the three
decorators the hot path touches:

  * ``Ingredient.config(fn)``      - run ``fn`` once and keep its locals as the config
  * ``Ingredient.config_hook(fn)`` - ignored (validation hook only)
  * ``Ingredient.capture(fn)``     - fill parameters the caller left out from the config

Nothing under ``pemp_b200/`` may import this module.
"""
import functools
import inspect
import sys


class Ingredient:
    def __init__(self, path="", ingredients=(), **_unused):
        self.path = path
        self.ingredients = list(ingredients)
        self.cfg = {}

    # -- decorators ---------------------------------------------------------
    def config(self, fn):
        harvested = {}

        def tracer(frame, event, arg):
            if frame.f_code is fn.__code__ and event == "return":
                harvested.update(frame.f_locals)
            return tracer

        previous = sys.gettrace()
        sys.settrace(tracer)
        try:
            fn()
        finally:
            sys.settrace(previous)
        self.cfg.update({k: v for k, v in harvested.items() if not k.startswith("_")})
        return fn

    def config_hook(self, fn):
        return fn

    def named_config(self, fn):
        return fn

    def capture(self, fn):
        sig = inspect.signature(fn)

        @functools.wraps(fn)
        def filled(*args, **kwargs):
            bound = sig.bind_partial(*args, **kwargs)
            for name in sig.parameters:
                if name not in bound.arguments and name in self.cfg:
                    kwargs[name] = self.cfg[name]
            return fn(*args, **kwargs)

        return filled

    # Experiment-only API used by entry files; present so imports do not fail.
    command = capture
    automain = main = staticmethod(lambda fn: fn)


Experiment = Ingredient
