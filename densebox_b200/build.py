"""Build the sm_100a shared library in-tree (nvcc cross-compiles without a GPU).

    python -m densebox_b200.build [--force]

Produces densebox_b200/csrc/libdensebox_b200.so. The library links cudart statically and resolves the one driver
entry point it needs (cuTensorMapEncodeTiled) at run time, so it loads on a box without a CUDA driver too.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libdensebox_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "densebox_b200.h")]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and p.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False, out=None, extra_flags=()):
    """out / extra_flags: build a VARIANT of the library somewhere else (tools/ab A/B measurements)."""
    variant = out is not None
    lib = out or LIB
    stamp = os.path.join(CSRC, ".build_stamp")
    dig = _digest()
    if not variant and not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    objdir = os.path.join(CSRC, "build", os.path.basename(lib) + ".obj") if variant else os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)

    def cc(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        cmd = [NVCC] + FLAGS + list(extra_flags) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(cc, _sources()))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    objs = [o for o, _ in results]
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-lpthread", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if not variant:
        with open(stamp, "w") as f:
            f.write(dig)
    return lib


if __name__ == "__main__":
    # python -m densebox_b200.build [--force] [--out PATH] [-DNAME=VALUE ...]
    _out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out,
                extra_flags=[a for a in sys.argv[1:] if a.startswith("-D")]))
