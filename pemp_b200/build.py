"""Build `pemp_b200/libpemp_b200.so` (the C-ABI library declared in `include/pemp_b200.h`) in-tree.

    python -m pemp_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpemp_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(LIB):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "pemp_b200.h")]
    return any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps)


def build(force=False, verbose=False, out=None, extra=None, only=None):
    """Compile every csrc/*.cu to an object (cached per flag set under build/obj) and link the shared library.
    `out` / `extra` build a variant with other -D flags for experiments (load it with PEMP_B200_LIB=<path>);
    `only` restricts the extra flags to the named source files (the rest reuse the default objects)."""
    import hashlib
    from concurrent.futures import ThreadPoolExecutor
    lib = out or LIB
    extra = list(extra or []) + os.environ.get("PEMP_NVCC_FLAGS", "").split()      # developer knob, e.g. -DPEMP_MPA_U=4
    if not force and out is None and not extra and not stale():
        return lib
    base_flags = [f for f in FLAGS if f not in ("-shared",)] + (["-Xptxas", "-v"] if verbose else [])

    def flags_for(src):
        return base_flags + (extra if only is None or os.path.basename(src) in only else [])

    def objdir_for(src):
        tag = hashlib.sha1(" ".join(flags_for(src)).encode()).hexdigest()[:10]
        d = os.path.join(HERE, "..", "build", "obj", tag)
        os.makedirs(d, exist_ok=True)
        return d
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "pemp_b200.h")]
    newest_hdr = max(os.path.getmtime(h) for h in hdrs)
    log = []

    def one(src):
        obj = os.path.join(objdir_for(src), os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_hdr) and not verbose:
            return obj
        cmd = [NVCC] + flags_for(src) + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        log.append(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(one, sources()))
    cmd = [NVCC] + FLAGS + objs + ["-o", lib]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print("".join(log))
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
