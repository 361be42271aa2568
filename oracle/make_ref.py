#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's own hot-path sources, UNMODIFIED, placed where the GPU box can run them.

    python oracle/make_ref.py            # in the build container, where /root/reference exists

The reference (joeyz0z/ConZIC) is a directory of Python scripts with no package metadata, so it cannot be pip
installed into baseline/_ref; this script is the committed recipe the task statement asks for instead.  It copies
the files of the path (gen_utils.py, control_gen_utils.py, utils.py, clip/clip.py, POS_classifier.py and the stop
list) byte for byte into oracle/_ref/ and writes their sha256 into MANIFEST.json.  oracle/_ref/ is git-ignored
(nothing of the reference enters the history) but travels to the GPU box with the snapshot, where
`bench.py --impl reference` times the reference's unmodified `generate_caption` on the host cores
(`cpu_baseline.kind == "reference"`).  Test infrastructure only: nothing under conzic_b200/ imports it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["gen_utils.py", "control_gen_utils.py", "utils.py", os.path.join("clip", "clip.py"), "POS_classifier.py",
         "stop_words.txt", "LICENSE"]


def make(src="/root/reference", quiet=False):
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src} does not exist (only the build container has the reference)")
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "sha256": manifest}, fh, indent=1)
    if not quiet:
        print(f"oracle/_ref: {len(FILES)} files copied unmodified from {src}")
    return DST


if __name__ == "__main__":
    make(*sys.argv[1:2])
