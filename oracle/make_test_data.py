#!/usr/bin/env python
"""Trimmed copies of the bundled ml-100k atomic files for the end-to-end tests (tests/data/ml-100k/): only the columns
the FOCF config loads (inter: user_id, item_id, rating; user: user_id, gender; item: item_id), same row order and the
same `name:type` headers, so ids / shuffles / splits come out exactly as from the originals.
TEST INFRASTRUCTURE ONLY.  Source: /root/reference/recbole/dataset_example/ml-100k/ (MovieLens-100k data files)."""
import os

SRC = "/root/reference/recbole/dataset_example/ml-100k"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "data", "ml-100k")


def trim(name, keep):
    with open(os.path.join(SRC, name)) as f:
        header = f.readline().rstrip("\n").split("\t")
        idx = [i for i, h in enumerate(header) if h.split(":")[0] in keep]
        lines = ["\t".join(header[i] for i in idx)]
        for line in f:
            parts = line.rstrip("\n").split("\t")
            lines.append("\t".join(parts[i] for i in idx))
    with open(os.path.join(DST, name), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    trim("ml-100k.inter", ("user_id", "item_id", "rating"))
    trim("ml-100k.user", ("user_id", "gender"))
    trim("ml-100k.item", ("item_id",))
    print({n: os.path.getsize(os.path.join(DST, n)) for n in os.listdir(DST)})
