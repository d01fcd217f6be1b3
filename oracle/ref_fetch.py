"""Recipe: stage the UNMODIFIED reference modules for bench.py's reference arm (test infrastructure, not product).

    python -m oracle.ref_fetch           (also run by __graft_entry__.build())

Copies /root/reference/code/{constants,utils,load,models}.py byte for byte into oracle/_ref/ -- a directory that is
git-ignored (the reference's sources never enter this repository's history) but NOT gpurun-ignored, so it travels
to the GPU box like the built .so files.  /root/reference exists only in the build container; on the GPU box the
staged copy is what `bench.py --impl reference` imports (through oracle/refrun.py's shims).  Nothing under
contrastiveprosthetics_b200/ reads this directory.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/code"
REF_DST = os.path.join(HERE, "_ref")
MODULES = ("constants", "utils", "load", "models")


def fetch(verbose=False):
    """Returns True when oracle/_ref holds the four modules (freshly copied or already staged)."""
    if not os.path.isdir(REF_SRC):
        return staged()
    os.makedirs(REF_DST, exist_ok=True)
    digest = {}
    for m in MODULES:
        src, dst = os.path.join(REF_SRC, m + ".py"), os.path.join(REF_DST, m + ".py")
        shutil.copyfile(src, dst)
        digest[m] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": REF_SRC, "sha256": digest}, open(os.path.join(REF_DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print("oracle/_ref: staged", ", ".join(m + ".py" for m in MODULES))
    return True


def staged():
    return all(os.path.isfile(os.path.join(REF_DST, m + ".py")) for m in MODULES)


if __name__ == "__main__":
    print("staged" if fetch(verbose=True) else "reference not available")
