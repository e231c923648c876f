import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The tests bind the in-tree C-ABI library; build it once if this checkout has not been built yet (nvcc cross-
    compiles without a GPU; the product itself never builds or falls back on its own)."""
    from pemp_b200 import build as _build
    if _build.stale():
        _build.build()


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def unpack_bits(packed, shape):
    n = int(np.prod(shape))
    return np.unpackbits(packed)[:n].reshape(shape)


def nrel(a, b):
    """Norm-wise relative error max|a-b| / max|b| (the tolerance semantics of SURVEY 7, hard part 1)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def screened_episodes(workload, spec, B, start=0):
    """The first B episode indices >= start whose smallest decision margin |fg - bg| through the reference restatement is
    >= 1e-5 (SURVEY 7 hard part 2, tier T2), their stacked batch, and how many candidates were rejected on the way.  For the
    bench-size streams the committed table (`pemp_b200/episode_screen.json`) answers; other specs are screened here."""
    from oracle import screen
    from pemp_b200 import episodes as E
    key = E.screen_key(workload, spec)
    if key in E.screen_table():
        acc = [i for i in E.screened_indices(workload, spec) if i >= start][:B]
        assert len(acc) == B
        return E.make_batch(spec, acc), acc, None
    acc, rejected, i = [], 0, start
    while len(acc) < B:
        m, _ = screen.episode_margin(workload, spec, i)
        if m >= screen.THRESHOLD:
            acc.append(i)
        else:
            rejected += 1
        i += 1
    return E.make_batch(spec, acc), acc, rejected
