"""Recipe for oracle/_ref: a byte-for-byte copy of the reference's package, taken where it lies.

    python oracle/make_ref.py            (dev container only: needs /root/reference)

The reference (amtsyplov/diversity-recommendations) is pure Python, so "building" it is copying
/root/reference/divrec to oracle/_ref/divrec.  oracle/_ref/ is git-ignored (no reference source enters the history) but
NOT gpurun-ignored, so the copy travels to the GPU box, where /root/reference does not exist.  It is used by
oracle/ref_run.py only: the literal-reference CPU arm of bench.py (`--impl reference`) and the check that the oracle
restatement agrees with the reference.  TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/divrec"
DST = os.path.join(HERE, "_ref", "divrec")


def make_ref(force: bool = False) -> str:
    if not os.path.isdir(SRC):
        if os.path.isdir(DST):
            return DST  # GPU box: the copy made in the dev container travelled with the snapshot
        raise FileNotFoundError(f"{SRC} is missing and no copy exists at {DST}")
    if os.path.isdir(DST) and not force:
        same = all(os.path.exists(os.path.join(DST, os.path.relpath(os.path.join(r, f), SRC)))
                   for r, _d, fs in os.walk(SRC) for f in fs if f.endswith(".py"))
        if same:
            return DST
    shutil.rmtree(os.path.dirname(DST), ignore_errors=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(os.path.dirname(DST), "README"), "w") as f:
        f.write("Verbatim copy of /root/reference/divrec made by oracle/make_ref.py. Not part of the repository history.\n")
    return DST


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
