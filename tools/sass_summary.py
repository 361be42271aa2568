#!/usr/bin/env python
"""Opcode histogram per kernel of libconzic.so from `cuobjdump -sass` (run here, no GPU needed): the evidence that the
GEMMs are tcgen05 / TMEM / TMA kernels (UTCHMMA, LDTM / STTM, UTMALDG, UTCBAR ...), the attention kernel mma.sync (HMMA),
the top-k a cluster kernel (UCGABAR / distributed shared memory accesses).

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "conzic_b200", "libconzic.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTMACCTL", "UTMACMDFLUSH", "LDTM", "STTM", "UTCATOM", "HMMA", "LDSM", "SYNCS",
       "UCGABAR", "CGABAR", "MUFU", "ATOMS", "RED", "SHFL", "LDG", "STG", "LDS", "STS", "BAR", "MEMBAR", "ERRBAR", "STL", "LDL")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, hist = None, collections.OrderedDict()
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            hist[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Za-z0-9_]+)*)", ln)
        if m and fn:
            hist[fn][m.group(1)] += 1
            full = m.group(1) + m.group(2)
            if m.group(1) in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTCBAR", "LDTM", "STTM", "HMMA"):
                hist[fn]["~" + full] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instructions per kernel, selected opcodes, and the distinct forms of the")
    print("# tensor / TMA / TMEM instructions (sm_100a)")
    for (fn, h), name in zip(hist.items(), demangle):
        short = re.sub(r"conzic::\(anonymous namespace\)::", "", name)
        short = re.sub(r"\(.*", "", short)
        total = sum(v for k, v in h.items() if not k.startswith("~"))
        keys = " ".join(f"{k}={h[k]}" for k in KEY if h.get(k))
        forms = " ".join(sorted(k[1:] for k in h if k.startswith("~")))
        print(f"{short}\n    total={total} {keys}\n    {forms}")


if __name__ == "__main__":
    main()
