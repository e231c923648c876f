"""Build variant libraries of libpemp_b200.so with extra -D flags on one source file (kernel experiments).
    python tools/build_variants.py mpa_tma.cu name1:-DX=1,-DY name2:-DZ ...   -> variants/libpemp_<name>.so
variants/ is git-ignored but travels to the GPU box; run with PEMP_B200_LIB=<path>."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pemp_b200 import build as B  # noqa: E402

src = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
outdir = os.path.join(root, "variants")
os.makedirs(outdir, exist_ok=True)


def one(spec):
    name, _, flags = spec.partition(":")
    out = os.path.join(outdir, f"libpemp_{name}.so")
    B.build(out=out, extra=[f for f in flags.split(",") if f], only=(src,))
    return out


B.build()   # default objects first (shared by every variant)
with ThreadPoolExecutor(max_workers=4) as ex:
    for o in ex.map(one, sys.argv[2:]):
        print(o)
