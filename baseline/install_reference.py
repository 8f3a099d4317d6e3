"""Re-host the reference for the `--impl reference` arm.

`pip install --target baseline/_ref /root/reference` is not possible: the reference
has no setup.py/pyproject ("Directory '/root/reference' is not installable"), and its
dependencies (grace_dl, cupy, pybloomfilter, mmh3, dahuffman) plus its precomputed
hash-table file are absent offline.  So the install is a byte-identical copy of
`pytorch/deepreduce.py` into the git-ignored `baseline/_ref/` (it travels to the GPU
box with gpurun); the missing third-party modules are provided by the minimal shims
in `baseline/shims/` (GRACE contract per SURVEY Appendix A; cupy packbits/unpackbits
via torch ops).  Nothing of deepreduce_b200 is on that path.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/pytorch/deepreduce.py"
DST_DIR = os.path.join(HERE, "_ref", "deepreduce_ref")
DST = os.path.join(DST_DIR, "deepreduce.py")


def install(verbose=True):
    if os.path.exists(DST):
        return DST
    if not os.path.exists(SRC):
        return None
    os.makedirs(DST_DIR, exist_ok=True)
    shutil.copyfile(SRC, DST)
    open(os.path.join(DST_DIR, "__init__.py"), "w").close()
    digest = hashlib.sha256(open(DST, "rb").read()).hexdigest()
    with open(os.path.join(DST_DIR, "SOURCE.txt"), "w") as f:
        f.write(f"copied unmodified from {SRC}\nsha256 {digest}\n")
    if verbose:
        print(f"[baseline] installed reference -> {DST} (sha256 {digest[:16]}…)")
    return DST


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
