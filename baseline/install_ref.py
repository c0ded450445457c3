#!/usr/bin/env python
"""
Installs the UNMODIFIED reference hot-path sources into baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU
box with the snapshot, where /root/reference does not exist).

    python baseline/install_ref.py            # copies /root/reference/Code -> baseline/_ref/Code, byte for byte

The prescribed `python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
/root/reference` was tried first and fails: the reference has neither setup.py nor pyproject.toml ("Directory '/root/reference'
is not installable") -- it is a flat directory of scripts imported by module name (Code/main.py:19-27).  The copy below is the
equivalent install: nothing is edited, every file is verified against the source by SHA-256 and the digests are written to
baseline/_ref/MANIFEST.json.  bench.py --impl reference imports these files through the reference's own public entry point
(Code/Test_Train.py:22 Training); nothing from pde_learn_b200/ or oracle/ is on that path.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PDL_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def main():
    code = os.path.join(SRC, "Code")
    if not os.path.isfile(os.path.join(code, "Loss.py")):
        print("reference not found at %s; baseline/_ref left as it is" % SRC)
        return 0 if os.path.isfile(os.path.join(DST, "Code", "Loss.py")) else 1
    out = os.path.join(DST, "Code")
    if os.path.isdir(out):
        shutil.rmtree(out)
    manifest = {}
    for root, _dirs, files in os.walk(code):
        for name in files:
            if not name.endswith(".py"):
                continue
            s = os.path.join(root, name)
            rel = os.path.relpath(s, SRC)
            d = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            assert sha(s) == sha(d)
            manifest[rel] = sha(d)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    print("installed %d reference files into %s" % (len(manifest), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
