#!/usr/bin/env python
"""Recipe for oracle/_ref: the UNMODIFIED reference package, placed where the GPU box can import it.

The reference (TangJiakai/RecBole-FairRec) is pure Python, so "building" it is copying its `recbole/` package tree from
/root/reference (which exists only in the build container) to oracle/_ref/recbole.  oracle/_ref/ is git-ignored (no
reference source enters the history) but NOT gpurun-ignored, so it travels to the GPU box with the snapshot, where
`bench.py --impl reference` and the drop-in tests import it through oracle/ref_shim/shim.py (kind = "reference").
Run by __graft_entry__.build() whenever /root/reference is present.  TEST / MEASUREMENT INFRASTRUCTURE ONLY."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("FAIRREC_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def main():
    src = os.path.join(SRC, "recbole")
    if not os.path.isdir(src):
        print(f"make_ref: {src} not found (nothing to do; an existing oracle/_ref is kept)")
        return 0
    dst = os.path.join(DST, "recbole")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    keep = (".py", ".yaml", ".yml")
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "dataset_example")]
        rel = os.path.relpath(root, src)
        for f in files:
            if f.endswith(keep):
                os.makedirs(os.path.join(dst, rel), exist_ok=True)
                shutil.copy2(os.path.join(root, f), os.path.join(dst, rel, f))
                n += 1
    with open(os.path.join(DST, "README"), "w") as fh:
        fh.write("Copy of /root/reference/recbole made by oracle/make_ref.py (git-ignored; unmodified reference code).\n")
    print(f"make_ref: {n} files -> {dst}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
