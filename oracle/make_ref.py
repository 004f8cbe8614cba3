"""Recipe for oracle/_ref/: an UNMODIFIED copy of the reference's pure-Python transform path, so that the reference itself
(not only the oracle port) can be timed on the GPU box, where /root/reference does not exist.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

Copies /root/reference/qsft/*.py and /root/reference/synt_exp/synt_src/*.py byte for byte into oracle/_ref/ (git-ignored:
the reference's sources never enter this repository's history; the directory travels to the GPU box with the snapshot like
the built .so files) and writes MANIFEST.json with their SHA-256.  TEST / BENCH INFRASTRUCTURE ONLY: imported by
bench.py's reference arm / cpu_baseline leg through oracle/ref_shim.py, never by the product."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
PACKAGES = ["qsft", os.path.join("synt_exp", "synt_src")]


def make(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: oracle/_ref left as it is", file=sys.stderr)
        return False
    manifest = {}
    for pkg in PACKAGES:
        src_dir, dst_dir = os.path.join(SRC, pkg), os.path.join(DST, pkg)
        os.makedirs(dst_dir, exist_ok=True)
        for name in sorted(os.listdir(src_dir)):
            if not name.endswith(".py"):
                continue
            shutil.copyfile(os.path.join(src_dir, name), os.path.join(dst_dir, name))
            manifest[os.path.join(pkg, name)] = hashlib.sha256(open(os.path.join(dst_dir, name), "rb").read()).hexdigest()
    # synt_exp is a namespace directory in the reference (no __init__.py): importable as is with oracle/_ref on sys.path
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} files copied unmodified from {SRC}")
    return True


if __name__ == "__main__":
    make()
