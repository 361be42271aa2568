#!/usr/bin/env python
"""Source-level stall attribution from an `ncu --set full --import-source on` capture (run here, no GPU needed):
warp-stall samples per CUDA source line, summed over the SASS instructions of that line, per captured launch.

    python tools/ncu_stalls.py gpurun_out/r01k/prof_gemm.ncu-rep [--top 15] [--launch 2]
"""
import collections
import csv
import io
import os
import subprocess
import sys

NAMES = ("stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_mio", "stall_barrier", "stall_membar",
         "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_dispatch", "stall_no_inst", "stall_sleep")


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 15
    only = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    launches, cur_file, first_file, hdr = [], None, None, None
    i = 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "File Path":
            cur_file = os.path.basename(r[1])
            if first_file is None:
                first_file = cur_file
            if cur_file == first_file:
                launches.append(collections.defaultdict(collections.Counter))
                launches[-1]["__fn__"]["name"] = 0
        elif r and r[0] == "Function Name":
            launches[-1]["__fn__"] = r[1]
        elif r and r[0] == "Line No":
            hdr = r
            si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
            cols = {n: hdr.index(n) for n in NAMES if n in hdr}
            line = None
            i += 1
            while i < len(rows) and rows[i] and rows[i][0] not in ("File Path", "Function Name", "Line No", "Kernel Name"):
                q = rows[i]
                if q[0].isdigit():
                    line = (cur_file, int(q[0]), q[1].strip()[:86])
                if line is not None and len(q) > si:
                    a = launches[-1][line]
                    try:
                        a["samples"] += int(q[si])
                        a["inst"] += int(q[ie])
                        for n, c in cols.items():
                            a[n] += int(q[c])
                    except ValueError:
                        pass
                i += 1
            continue
        i += 1
    for li, agg in enumerate(launches):
        if only is not None and li != only:
            continue
        fn = agg.pop("__fn__")
        tot = sum(a["samples"] for a in agg.values()) or 1
        kinds = collections.Counter()
        for a in agg.values():
            for n in NAMES:
                kinds[n] += a[n]
        print(f"== launch {li}: {str(fn)[:110]}\n   samples {tot}; by reason (%): " +
              ", ".join(f"{n[6:]} {100 * v / tot:.1f}" for n, v in kinds.most_common(8)))
        for (f, ln, src), a in sorted(agg.items(), key=lambda x: -x[1]["samples"])[:top]:
            main_reason = max(NAMES, key=lambda n: a[n])
            print(f"  {100 * a['samples'] / tot:5.1f}%  {f}:{ln:<5d} {src:86s} [{main_reason[6:]} {a[main_reason]}; inst {a['inst']}]")


if __name__ == "__main__":
    main()
