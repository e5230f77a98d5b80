"""Build libdivrec_b200.so in-tree with nvcc for sm_100a (no other architecture is supported).

    python diversity-recommendations_b200/build.py [--force]

The shared library is the C-ABI boundary declared in include/divrec_b200.h.  It is git-ignored
but travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# experiment builds (tools/ab.py): DIVREC_BUILD_TAG=<tag> DIVREC_BUILD_DEFS="-DX -DY" -> libdivrec_b200.<tag>.so, objects in build.<tag>/
_TAG = os.environ.get("DIVREC_BUILD_TAG", "")
OUT = os.path.join(HERE, f"libdivrec_b200.{_TAG}.so" if _TAG else "libdivrec_b200.so")
OBJ_DIR = os.path.join(HERE, f"build.{_TAG}" if _TAG else "build")
SOURCES = ["api.cu", "score_filter.cu", "select_topk.cu", "rescue.cu", "prep.cu", "metrics.cu", "bpr.cu", "mmr.cu", "score_filter2.cu", "csr.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--use_fast_math=false",
]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"] + os.environ.get("DIVREC_BUILD_DEFS", "").split()


def _digest():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    with open(os.path.join(HERE, "..", "include", "divrec_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "stamp")
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        with open(os.path.join(OBJ_DIR, src + ".ptxas.log"), "w") as f:
            f.write(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
