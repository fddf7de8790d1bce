"""Recipe for ``oracle/_ref``: the UNMODIFIED reference (crystal22/C2DSR), made importable on the GPU box.

    python oracle/make_ref.py            # needs /root/reference (present in the build container only)

The reference is pure Python (no native sources, no setup.py: nothing to compile or pip-install), so "building" it
means placing its hot-path modules -- main.py, trainer.py, dataloader.py, models/, utils/ -- under the git-ignored
``oracle/_ref/`` (listed in .gitignore, NOT in .gpurunignore: it travels to the GPU box like the built .so).  Nothing
from it is committed, and no product code imports it: ``bench.py --impl reference`` / ``cpu_baseline`` time it on the
box's host cores (``kind: "reference"``) and tests/ may compare the oracle with it.  ``__graft_entry__.build()`` runs
this when /root/reference exists; on the GPU box the prebuilt copy is used as is.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("C2DSR_REFERENCE", "/root/reference")
WANTED = ("main.py", "trainer.py", "dataloader.py", "models", "utils")


def make(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: keeping {'the existing' if os.path.isdir(DST) else 'no'} oracle/_ref")
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    for name in WANTED:
        s, d = os.path.join(SRC, name), os.path.join(DST, name)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
    if verbose:
        print(f"oracle/_ref <- {SRC} ({', '.join(WANTED)})")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
