"""Recipe for oracle/_ref/: the UNMODIFIED reference's Python packages for the sampling path, staged where they can
travel to the GPU box (oracle/_ref/ is git-ignored -- no reference source enters the history -- but not
gpurun-ignored).  Run in the build container, where /root/reference exists:

    python oracle/make_ref.py            (also called by __graft_entry__.build())

`pip install --target baseline/_ref /root/reference` is not an option: the reference's packages carry no
__init__.py, so its setup.py (find_packages()) builds an EMPTY wheel (tried; outcome recorded in DESIGN.md section 7).
The files are therefore copied verbatim: model/ (LFAE + BaseDM_adaptor) and config/DM/*.yaml -- what
scripts/DM/valid.py imports for sampling.  Consumers: bench.py's reference legs (`--impl reference`, `cpu_baseline`,
`gpu_eager_reference`) through oracle/ref_shims.py.  Test infrastructure, never the product.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("EXTDM_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def make(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "model")):
        return None                                   # GPU box: only the pre-staged copy is used
    for sub in ("model", os.path.join("config", "DM")):
        dst = os.path.join(DST, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, sub), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(SRC + "\n")
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        sys.stderr.write(f"oracle/_ref: {n} files staged from {SRC}\n")
    return DST


if __name__ == "__main__":
    print(make(verbose=True))
