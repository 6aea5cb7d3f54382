"""Recipe for oracle/_ref/: makes the UNMODIFIED reference model travel to the GPU box.

    python -m oracle.build_ref          (run by __graft_entry__.build() whenever /root/reference exists)

The reference is pure Python (no native code to compile): "building" it means copying its model file,
byte for byte, from where it lies under /root/reference into oracle/_ref/, which is git-ignored (never
part of this repo's history) but NOT gpurun-ignored, so the copy rides along with the built .so files.
oracle/ref_shim.py then imports it under the installed transformers (runtime shims only, file untouched)
and bench.py's `--impl reference` / `--impl reference-gpu` arms and cpu_baseline leg time THE REFERENCE
ITSELF (cpu_baseline.kind = "reference") instead of the oracle port.  A sha256 of the source is written
next to it and checked on load, so a stale or edited copy is refused.

TEST / BENCH INFRASTRUCTURE ONLY: nothing under ergm_b200/ reads oracle/_ref/.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TREE = "/root/reference"
OUT = os.path.join(HERE, "_ref")
FILES = ("src/model.py",)


def build(verbose=True):
    if not os.path.isdir(REF_TREE):
        if verbose:
            print("oracle/build_ref: %s absent (GPU box): using the prebuilt oracle/_ref" % REF_TREE)
        return os.path.isfile(os.path.join(OUT, "model.py"))
    os.makedirs(OUT, exist_ok=True)
    sums = []
    for rel in FILES:
        src = os.path.join(REF_TREE, rel)
        dst = os.path.join(OUT, os.path.basename(rel))
        shutil.copyfile(src, dst)
        sums.append("%s  %s" % (hashlib.sha256(open(dst, "rb").read()).hexdigest(), os.path.basename(rel)))
    open(os.path.join(OUT, "SHA256SUMS"), "w").write("\n".join(sums) + "\n")
    if verbose:
        print("oracle/build_ref: copied %s -> %s" % (", ".join(FILES), OUT))
    return True


def verify():
    """True when oracle/_ref/model.py exists and matches its recorded checksum."""
    sums = os.path.join(OUT, "SHA256SUMS")
    if not os.path.isfile(sums):
        return False
    for line in open(sums).read().split("\n"):
        if not line.strip():
            continue
        digest, name = line.split()
        p = os.path.join(OUT, name)
        if not os.path.isfile(p) or hashlib.sha256(open(p, "rb").read()).hexdigest() != digest:
            return False
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
