"""Recipe for oracle/_ref/: the UNMODIFIED reference module, staged so that it can travel to the GPU box.

TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as oracle/macaw_oracle.py: only tests/, __graft_entry__ and bench.py's
CPU legs may use it).

The reference (lyuchenyang/Macaw-LLM) is pure Python: there is nothing to compile.  "Building" the reference arm means
copying the ONE file the hot path lives in — /root/reference/modeling.py — byte for byte into oracle/_ref/ (git-ignored,
so it never enters the history; NOT gpurun-ignored, so it ships with the snapshot like a built .so).  The two
transformers-5.x compatibility shims of SURVEY.md §8c are applied at import time by oracle/ref_runner.py; the file
itself is not edited (its sha256 is recorded and re-checked on load).

  python oracle/make_ref.py            # no-op when /root/reference is absent (GPU box: uses the staged copy)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SRC = "/root/reference/modeling.py"


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def build(verbose: bool = False) -> bool:
    """Stage the reference module.  Returns True when oracle/_ref/modeling.py is present afterwards."""
    dst = os.path.join(REF_DIR, "modeling.py")
    if os.path.exists(SRC):
        os.makedirs(REF_DIR, exist_ok=True)
        if not os.path.exists(dst) or sha256(dst) != sha256(SRC):
            shutil.copyfile(SRC, dst)
        with open(os.path.join(REF_DIR, "SOURCE.json"), "w") as f:
            json.dump({"source": SRC, "sha256": sha256(dst), "edited": False,
                       "note": "verbatim copy of the reference hot-path module; shims are applied at import time"}, f)
        if verbose:
            print(f"[make_ref] staged {SRC} -> {dst}")
    return os.path.exists(dst)


if __name__ == "__main__":
    print("oracle/_ref ready" if build(verbose=True) else "oracle/_ref absent (no /root/reference here)")
