#!/usr/bin/env python
"""captions/sec of the Gibbs-BERT caption-polishing path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: a full generate_caption call on 64 synthetic images per
GPU (BASELINE config 2: sequential order, sentence_len 10, candidate_k 200, 5 sweeps = 50 Gibbs steps, image
encoding included, model load excluded).  Images are independent, so N GPUs process N x 64 images (weak
scaling) with one NCCL all-gather of the final ids/scores per call.

  value        captions/s with the pixel tensors already resident in HBM (device loop, no host reads);
  e2e          captions/s through the public drop-in API conzic_b200.gen_utils.generate_caption with the pixels
               in pinned HOST memory: H2D copy, per-sweep D2H reads of ids/scores and string decoding included;
  roofline     the dominant kernel (tcgen05 GEMM of the CLIP/BERT towers): executed FLOPs / CUDA-event time of
               its launches during one extra profiled step, against the measured sustained bf16 peak;
  cpu_baseline the CPU oracle (a torch-CPU restatement of the reference loop pinned against the unmodified
               reference) on a bounded sample, all host cores;
  --impl reference   the same CPU path as its own arm (the Python reference cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from conzic_b200 import dist as cdist  # noqa: E402
from conzic_b200 import synth  # noqa: E402

N_LEN, TOP_K, SWEEPS, BATCH = 10, 200, 5, 64
TEMP, ALPHA, BETA = 0.1, 0.02, 2.0
METRIC = "captions/sec (len=10, top_k=200, 5 iters)"
# algorithmic FLOPs of one caption at n=10, K=200, 5 sweeps
ALGO_TFLOP_PER_CAPTION = 10.806


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


def ncu_traffic():
    """dram read+write bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/traffic.json, written by tools/ncu_summary.py); None when no capture has been summarised."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    return d.get("dram_bytes_per_launch"), {"kernel": d.get("kernel"), "shape": d.get("shape"),
                                            "algorithmic_bytes_per_launch": d.get("algorithmic_bytes_per_launch"),
                                            "source": d.get("source")}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy or sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference loop)
# ------------------------------------------------------------------------------------------------------
def cpu_sample(B=2, threads=None):
    """One bounded sample of the workload on the host cores: first sweep (growing captions) + one full-length
    sweep for B images; captions/s for 5 sweeps = B / (t_first + 4 * t_full).  Returns (captions/s, seconds)."""
    from oracle import conzic_oracle as orc
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    global _CPU_ORACLE
    if "_CPU_ORACLE" not in globals():
        o = orc.Oracle(synth.make_bert_state_dict(0), synth.make_clip_state_dict(0, vision=True),
                       synth.SynthBertTokenizer(), synth.SynthCLIPTokenizer(), full_logits=True)
        _CPU_ORACLE = o
    o = _CPU_ORACLE
    pix = torch.stack([synth.make_pixel_values(i) for i in range(B)])
    tm = synth.make_token_mask()
    with torch.no_grad():
        t0 = time.perf_counter()
        o.generate(pix, tm, synth.SYNTH_PROMPT, order="sequential", max_len=N_LEN, top_k=TOP_K, temperature=TEMP,
                   alpha=ALPHA, beta=BETA, max_iters=2)
        t = time.perf_counter() - t0
    # the two sweeps are timed together; split by CLIP token counts (105 vs 150 tokens per candidate)
    t_first, t_full = t * 105.0 / 255.0, t * 150.0 / 255.0
    return B / (t_first + (SWEEPS - 1) * t_full), t


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B = 1
    for _ in range(args.warmup):
        cpu_sample(B, cores)
    vals, secs = [], []
    for _ in range(args.steps):
        v, s = cpu_sample(B, cores)
        vals.append(v); secs.append(s)
    value = len(vals) / sum(1.0 / v for v in vals)  # harmonic mean = total captions / total time
    sample = (f"{B} image(s) x (first sweep + one full-length sweep) of the len=10/K=200 workload per step, "
              f"5-sweep time extrapolated as t_first + 4*t_full; CPU cost is linear in images (SURVEY.md 6)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "captions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sum(secs) / len(secs),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "captions/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": f"{BATCH}-image batch per GPU, sequential order, sentence_len {N_LEN}, candidate_k {TOP_K}, "
                        f"{SWEEPS} sweeps (BASELINE config 2), bert-base + CLIP ViT-B/32 shapes, synthetic weights",
            "images_per_gpu": BATCH, "n_gpus": n, "sharding": "images by global index, one all-gather per call",
            "l2": "per-step activations (>1 GB) exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
class Job:
    def __init__(self, rank, local_rank, world, precision):
        from conzic_b200 import runtime
        from conzic_b200.clip.clip import CLIP
        from conzic_b200.models import BertMLM
        self.rank, self.world = rank, world
        self.dev = torch.device("cuda", local_rank)
        torch.cuda.set_device(self.dev)
        os.environ["CONZIC_PRECISION"] = precision
        self.bert = BertMLM(synth.make_bert_state_dict(0))
        self.clip = CLIP(state_dict=synth.make_clip_state_dict(0), tokenizer=synth.SynthCLIPTokenizer(),
                         processor=synth.SynthProcessor()).to(self.dev)
        self.tok = synth.SynthBertTokenizer()
        self.eng = runtime.engine_for(self.bert, self.clip, self.tok, device=self.dev)
        idx = range(rank * BATCH, (rank + 1) * BATCH)  # images keyed by global index
        self.pix_host = torch.stack([synth.make_pixel_values(i) for i in idx]).pin_memory()
        self.pix_dev = self.pix_host.to(self.dev)
        self.names = [f"img{i}.jpg" for i in idx]
        self.logger = logging.getLogger("bench")
        self.logger.addHandler(logging.NullHandler())
        self.logger.propagate = False
        self.L = N_LEN + 5
        self.init_ids = torch.tensor([self.tok.encode(synth.SYNTH_PROMPT + "[MASK]" * N_LEN)] * BATCH, device=self.dev)
        self.clip_ref = torch.zeros(BATCH, device=self.dev)

    def device_step(self):
        """Hot path only, everything resident: image encode + 50 Gibbs steps + the result gather."""
        eng = self.eng
        img = self.clip.compute_image_representation_from_pixels(self.pix_dev)
        inp = self.init_ids.clone()
        tm = synth.make_token_mask(self.dev)
        holds = [True] * 4 + [False] * N_LEN + [False]
        holds[0] = False
        for _ in range(SWEEPS):
            for ii in range(N_LEN):
                pos = 4 + ii
                eng.gibbs_step(inp, tm, img, pos, ii == N_LEN - 1, TOP_K, TEMP, ALPHA, BETA, sum(holds[:pos]),
                               sum(holds[pos + 1:]), out_clip_ref=self.clip_ref)
                holds[pos] = True
        ids, sc = cdist.gather_ids_scores(inp.to(torch.int32), self.clip_ref, self.world)
        return ids, sc

    def api_step(self):
        """The call a user makes: host pixels in, caption strings out."""
        from conzic_b200 import gen_utils
        tm = synth.make_token_mask(self.dev)
        texts, scores = gen_utils.generate_caption(self.names, self.bert, self.clip, self.tok, self.pix_host, tm,
                                                   self.logger, prompt=synth.SYNTH_PROMPT, batch_size=BATCH,
                                                   max_len=N_LEN, top_k=TOP_K, temperature=TEMP, max_iter=SWEEPS,
                                                   alpha=ALPHA, beta=BETA, generate_order="sequential")
        if self.world > 1:  # rank 0 collects every rank's final captions (a few KB of strings)
            cdist.gather_objects(texts[-2], self.world)
        return texts, scores


def timed(fn, steps, world, dev):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    return cdist.max_over_ranks(a.elapsed_time(b), world, dev)


def run_ours(args, rank, local_rank, world):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    job = Job(rank, local_rank, world, args.precision)
    eng = job.eng
    for _ in range(max(args.warmup, 3)):
        job.device_step()
    l0 = eng.launch_count()
    with ClockSampler(local_rank) as cs:
        ms = timed(job.device_step, args.steps, world, job.dev)
    launches = (eng.launch_count() - l0) * world
    clocks = cs.summary()
    value = world * BATCH * args.steps / (ms / 1000.0)
    # end to end through the drop-in API (host pixels, strings out)
    job.api_step()
    t0 = time.perf_counter()
    if world > 1:
        torch.distributed.barrier()
    for _ in range(args.steps):
        job.api_step()
    torch.cuda.synchronize(job.dev)
    if world > 1:
        torch.distributed.barrier()
    e2e_s = cdist.max_over_ranks(time.perf_counter() - t0, world, job.dev)
    e2e = world * BATCH * args.steps / e2e_s
    h2d = job.pix_host.numel() * 4
    d2h = SWEEPS * (BATCH * job.L * 8 + BATCH * 4)
    # one extra profiled step: CUDA events around every launch, by category
    eng.profile(True)
    job.device_step()
    prof = eng.profile_read()
    eng.profile(False)
    pk = peaks()
    traffic, traffic_detail = ncu_traffic()
    g_ms, g_flops, g_n = prof["gemm"]
    achieved = g_flops / (g_ms / 1000.0) / 1e12 if g_ms > 0 else 0.0
    total_ms = sum(v[0] for v in prof.values())
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "captions/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.precision.replace("bf16x3", "bf16 (3-pass split)"),
                "data": "synthetic", "config": workload_config(world), "clocks": clocks,
                "e2e": {"value": e2e, "unit": "captions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches,
                "roofline": {"kernel": "gemm_persist_kernel<2,*> + gemm_wide_kernel (the CLIP towers' linears: CTA "
                                       "pairs, tcgen05 cta_group::2, TMA-fed); average over the launches of one step",
                             "bound": "tensor",
                             "achieved": achieved, "peak": pk["tf"], "unit": "TFLOP/s", "frac": achieved / pk["tf"],
                             "peak_source": f"{pk['src']} sustained bf16", "traffic": traffic,
                             "traffic_detail": traffic_detail,
                             "launches_profiled": g_n, "avg_launch_ms": g_ms / max(g_n, 1),
                             "flops_per_launch": g_flops / max(g_n, 1),
                             "share_of_step": g_ms / total_ms if total_ms else None},
                "step_breakdown_ms": {k: round(v[0], 3) for k, v in prof.items()},
                "algorithmic": {"tflop_per_caption": ALGO_TFLOP_PER_CAPTION,
                                "tflops_per_gpu": value / world * ALGO_TFLOP_PER_CAPTION,
                                "frac_of_peak": value / world * ALGO_TFLOP_PER_CAPTION / pk["tf"],
                                "note": "reference-algorithm FLOPs (every candidate encoded in full); the engine "
                                        "executes fewer because the caption prefix is encoded once per image"}}
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            cpu_sample(2, cores)  # warm-up (cold first call is several times slower)
            v, s = cpu_sample(2, cores)
            line["cpu_baseline"] = {"value": v, "unit": "captions/s", "cores": cores, "kind": "port",
                                    "sample": f"2 images x (first sweep + one full-length sweep), {s:.1f} s of CPU work, "
                                              "5-sweep time extrapolated as t_first + 4*t_full"}
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank, local_rank, world = cdist.env_world()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        cdist.init("nccl")
    run_ours(args, rank, local_rank, world)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
