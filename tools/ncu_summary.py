#!/usr/bin/env python
"""Summarise ncu captures (run here, no GPU needed) into small text/JSON files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/r01/launches_step.csv profiles/r01_launches_summary.txt
    python tools/ncu_summary.py full gpurun_out/r01/prof_gemm.ncu-rep profiles/r01_gemm_full.txt [--traffic M N K]
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("conzic::<unnamed>::", "").replace("unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as fh:
        fh.write(f"# per-kernel totals of {os.path.basename(src)} (ncu --metrics gpu__time_duration.sum, cold cache, serialised)\n")
        fh.write(f"# total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            fh.write(f"{k[:64]:64s} n={v[0]:5d} total_ms={v[1] / 1e6:9.3f} avg_us={v[1] / v[0] / 1e3:8.1f} share={v[1] / tot:.3f}\n")
    print(open(dst).read())


def full(src, dst, traffic_shape=None, pick=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none, {os.path.basename(src)}; one column per captured launch\n")
        for i, h in enumerate(hdr):
            if h in ("Kernel Name", "Grid Size", "Block Size") or h in KEYS:
                fh.write(f"{h} [{units[i]}]: " + " | ".join(r[i] for r in data) + "\n")
    print(open(dst).read())
    if traffic_shape:
        M, N, K = traffic_shape
        def col(name):
            i = hdr.index(name)
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
            return [float(r[i].replace(",", "")) * mult for r in data]
        rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
        if pick:  # only the captured launches that have this shape
            rd, wr = [rd[i] for i in pick], [wr[i] for i in pick]
        t = {"kernel": data[0][hdr.index("Kernel Name")].split("(")[0], "shape": {"M": M, "N": N, "K": K},
             "dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd),
             "algorithmic_bytes_per_launch": 2.0 * (M * K + N * K + M * N),
             "launches_averaged": pick if pick else list(range(len(rd))),
             "source": "profiles/" + os.path.basename(dst) + " (ncu --set full capture " + os.path.basename(src) + ")"}
        json.dump(t, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
        print(json.dumps(t))


def small(srcs, dst, peak_gbs=None):
    """One table row per capture: duration, DRAM bytes, achieved GB/s against the measured HBM peak -- for the
    HBM / latency bound kernels of the step (top-k, cosine logits, selection, assembly, LayerNorm, embeddings)."""
    if peak_gbs is None:
        p = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak_gbs = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6549.0
    lines = [f"| kernel | grid x block | regs | time us | DRAM read MB | DRAM write MB | achieved GB/s | % of HBM peak ({peak_gbs:.0f} GB/s) | "
             "L2 throughput % | SM throughput % | warps active % | capture |", "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for src in srcs:
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            def val(name, scale=True):
                if name not in hdr:
                    return float("nan")
                i = hdr.index(name)
                v = float(r[i].replace(",", "") or "nan")
                if scale:
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[i], 1.0)
                return v
            t_us = val("gpu__time_duration.sum")
            rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
            gbs = (rd + wr) / (t_us * 1e-6) / 1e9 if t_us > 0 else 0.0
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("conzic::<unnamed>::", "")
            lines.append(f"| {name} | {r[hdr.index('Grid Size')]} x {r[hdr.index('Block Size')]} | "
                         f"{val('launch__registers_per_thread', False):.0f} | {t_us:.1f} | {rd / 1e6:.2f} | {wr / 1e6:.2f} | {gbs:.0f} | "
                         f"{100 * gbs / peak_gbs:.1f} | {val('lts__throughput.avg.pct_of_peak_sustained_elapsed', False):.1f} | "
                         f"{val('sm__throughput.avg.pct_of_peak_sustained_elapsed', False):.1f} | "
                         f"{val('sm__warps_active.avg.pct_of_peak_sustained_active', False):.1f} | {os.path.basename(src)} |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "small":
        small(sys.argv[3:], sys.argv[2])
    else:
        shape = None
        if "--traffic" in sys.argv:
            i = sys.argv.index("--traffic")
            shape = tuple(int(x) for x in sys.argv[i + 1:i + 4])
        pick = None
        if "--pick" in sys.argv:
            pick = [int(x) for x in sys.argv[sys.argv.index("--pick") + 1].split(",")]
        full(sys.argv[2], sys.argv[3], shape, pick)
