"""Recipe that makes the UNMODIFIED reference importable where /root/reference does not exist (the GPU box).

    python -m oracle.vendor_reference        # also run by __graft_entry__.build() when /root/reference is present

Copies /root/reference/rbnet/*.py byte for byte into oracle/_ref/rbnet/ -- a build output: oracle/_ref/ is git-ignored
(no reference source ever enters this repository's history) but not gpurun-ignored, so it travels to the GPU box like a
built .so.  `bench.py --impl reference --workload c1` then times the literal reference (T0: its own Python loop nest,
its own autograd outside pass) on the box's host cores; oracle/reference_loader.py falls back to this copy when
/root/reference is absent.  A SHA-256 manifest is written beside the copy so that "unmodified" can be checked.

TEST / BASELINE INFRASTRUCTURE ONLY: nothing under rbn_b200/ imports oracle/.
"""
import hashlib
import json
import os
import shutil

SRC = "/root/reference/rbnet"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def vendor():
    if not os.path.isdir(SRC):
        return False
    out = os.path.join(DST, "rbnet")
    os.makedirs(out, exist_ok=True)
    manifest = {}
    for name in sorted(os.listdir(SRC)):
        if not name.endswith(".py"):
            continue
        shutil.copyfile(os.path.join(SRC, name), os.path.join(out, name))
        with open(os.path.join(out, name), "rb") as f:
            manifest[name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=SRC, note="byte-for-byte copy of the reference package; not part of this repository",
                       sha256=manifest), f, indent=1)
    return True


if __name__ == "__main__":
    print("vendored" if vendor() else f"{SRC} not present: nothing to do")
