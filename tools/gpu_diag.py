#!/usr/bin/env python
"""Bring-up diagnostics for the B200 box: each stage runs in its own process under a timeout, prints what it
measured and appends a JSON record to gpurun_out/diag.jsonl.  Not part of the product or of the test suite.

    python tools/gpu_diag.py                 # all stages
    python tools/gpu_diag.py --stage gemm    # one stage, in this process
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")
STAGES = ["gemm", "gemm_simt", "bert", "topk", "clip", "step", "perf_gemm", "perf_step"]


def emit(stage, rec):
    os.makedirs(OUT, exist_ok=True)
    rec = dict(stage=stage, **rec)
    print(json.dumps(rec), flush=True)
    with open(os.path.join(OUT, "diag.jsonl"), "a") as fh:
        fh.write(json.dumps(rec) + "\n")


def ref_linear(A, W, bias, resid, act, bf16_round):
    import torch
    if bf16_round:
        A, W = A.bfloat16().float(), W.bfloat16().float()
    y = A.double() @ W.double().t()
    if bias is not None:
        y = y + bias.double()
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    if resid is not None:
        y = y + resid.double()
    return y.float()


def stage_gemm(impl="tcgen05"):
    import torch
    import gpu_common as gc
    shapes = [(128, 128, 64, 0), (128, 128, 512, 0), (256, 256, 128, 0), (200, 512, 512, 1), (1000, 1536, 512, 0),
              (333, 2048, 512, 1), (4096, 512, 2048, 0), (64, 30522, 768, 0), (18, 768, 768, 2), (77, 2304, 768, 0)]
    if impl != "tcgen05":
        shapes = shapes[:6]
    for prec in ("bf16", "bf16x3"):
        eng = gc.engine(prec, impl)
        for (M, N, K, act) in shapes:
            g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
            A = torch.randn(M, K, device="cuda", generator=g)
            W = torch.randn(N, K, device="cuda", generator=g) * 0.05
            bias = torch.randn(N, device="cuda", generator=g)
            resid = torch.randn(M, N, device="cuda", generator=g) if N % 4 == 0 else None
            Npad = N
            if N % 4:
                # out ld must keep rows 16B aligned: pad W with zero rows like the engine pads the logits row
                Npad = (N + 3) & ~3
                W = torch.cat([W, torch.zeros(Npad - N, K, device="cuda")])
                bias = torch.cat([bias, torch.zeros(Npad - N, device="cuda")])
            out = eng.debug_linear(A, W, bias, resid, act)
            torch.cuda.synchronize()
            ref = ref_linear(A, W, bias, resid, act, prec == "bf16")
            err = float((out - ref).abs().max())
            scale = float(ref.abs().max())
            emit("gemm" if impl == "tcgen05" else "gemm_simt",
                 dict(prec=prec, M=M, N=N, K=K, act=act, max_abs_err=err, ref_absmax=scale,
                      ok=bool(err < (2e-3 if prec == "bf16" else 2e-4) * max(scale, 1.0))))


def stage_bert():
    import torch
    import gpu_common as gc
    from oracle import conzic_oracle as orc
    from conzic_b200 import synth
    sd = gc.weights("bert")
    inp = torch.tensor([[101, 3746, 1997, 1037, 103, 103, 103, 103, 102],
                        [101, 3746, 1997, 1037, 5000, 103, 7000, 2500, 102]])
    with torch.no_grad():
        ref = orc.bert_mlm_head(sd, orc.bert_encoder(sd, inp)[:, 5])
    for prec in ("bf16x3", "bf16"):
        for impl in ("tcgen05", "simt_debug"):
            try:
                eng = gc.engine(prec, impl)
                out = eng.bert_mlm_row(inp.cuda(), 5).cpu()
                err = float((out - ref).abs().max())
                emit("bert", dict(prec=prec, impl=impl, max_abs_err=err, ref_std=float(ref.std()),
                                  argmax_same=bool((out.argmax(1) == ref.argmax(1)).all())))
            except Exception as e:  # noqa: BLE001
                emit("bert", dict(prec=prec, impl=impl, error=str(e)[:300]))


def stage_topk():
    import torch
    import gpu_common as gc
    from conzic_b200 import synth
    eng = gc.engine("bf16x3", "simt_debug")
    torch.manual_seed(0)
    V = synth.BERT_VOCAB
    for name, logits in (("smooth", torch.randn(4, V) * 0.56), ("peaked", torch.randn(4, V) * 3.5)):
        mask = synth.make_token_mask()
        for K in (8, 200, 512, 1000):
            probs = torch.softmax(logits / 0.1, dim=-1) * mask
            rp, ri = probs.topk(K, dim=-1)
            ld = (V + 3) & ~3
            lg = torch.zeros(4, ld)
            lg[:, :V] = logits
            gp, gi = eng.topk_mask(lg.cuda()[:, :V], mask.cuda(), 0.1, K)
            gp, gi = gp.cpu(), gi.cpu()
            nz = rp > 0
            # where the reference probabilities are non-zero and distinct the ids must match exactly
            id_bad = int((gi[nz] != ri[nz]).sum())
            rel = float(((gp - rp).abs() / rp.clamp_min(1e-30))[nz].max()) if bool(nz.any()) else 0.0
            # tie contract: among zero-probability picks, indices ascending and the lowest available
            tie_ok = True
            for r in range(4):
                z = gi[r][~nz[r]]
                if z.numel() > 1 and not bool((z[1:] > z[:-1]).all()):
                    tie_ok = False
            desc = bool((gp[:, 1:] <= gp[:, :-1]).all())
            emit("topk", dict(case=name, K=K, id_mismatch_nonzero=id_bad, max_rel_prob_err=rel, n_zero=int((~nz).sum()),
                              tie_order_ok=tie_ok, sorted_desc=desc))


def stage_clip():
    import torch
    import gpu_common as gc
    from oracle import conzic_oracle as orc
    from conzic_b200 import synth
    sd = gc.weights("clip")
    torch.manual_seed(1)
    N, T = 37, 11
    ids = torch.randint(300, 40000, (N, T))
    ids[:, 0] = synth.CLIP_BOS
    lens = torch.randint(3, T + 1, (N,))
    for i in range(N):
        ids[i, lens[i] - 1:] = synth.CLIP_EOS
    with torch.no_grad():
        ref = orc.clip_text_embeds(sd, ids)
    for prec in ("bf16x3", "bf16"):
        for impl in ("tcgen05", "simt_debug"):
            try:
                eng = gc.engine(prec, impl)
                out = eng.clip_text_encode(ids.int().cuda()).cpu()
                err = float((out - ref).abs().max())
                cos = torch.nn.functional.cosine_similarity(out, ref, dim=-1)
                emit("clip", dict(prec=prec, impl=impl, max_abs_err=err, ref_absmax=float(ref.abs().max()),
                                  min_cos=float(cos.min())))
            except Exception as e:  # noqa: BLE001
                emit("clip", dict(prec=prec, impl=impl, error=str(e)[:300]))
    # similarity
    eng = gc.engine("bf16x3", "simt_debug")
    img = torch.randn(3, 512)
    txt = torch.randn(3 * 12, 512)
    rs, rr = orc.image_text_similarity(img, txt, torch.tensor(synth.LOGIT_SCALE))
    gs, gr = eng.image_text_similarity(img.cuda(), txt.cuda())
    emit("clip", dict(sim_score_err=float((gs.cpu() - rs).abs().max()), sim_ref_err=float((gr.cpu() - rr).abs().max())))


def stage_step():
    import gpu_common as gc
    for prec, impl in (("bf16x3", "tcgen05"), ("bf16", "tcgen05"), ("bf16x3", "simt_debug")):
        for name in ("seq_b2_n4_k8", "shuffle_b3_n5_k16_multi", "senti_shuffle_neg_b2_n4_k8", "peaked_seq_b2_n4_k32",
                     "random_b2_n3_k8", "senti_seq_b2_n4_k8", "seq_b1_n10_k200"):
            g = gc.load_golden(name)
            case = g["case"]
            try:
                eng = gc.engine(prec, impl, case.get("peaked", False), case.get("multi", False))
                m = gc.replay_fixture(eng, g)
                emit("step", dict(prec=prec, impl=impl, fixture=name, **m))
            except Exception as e:  # noqa: BLE001
                emit("step", dict(prec=prec, impl=impl, fixture=name, error=str(e)[:300]))
        gc.drop_engines()


def _time(fn, iters=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def stage_perf_gemm():
    import torch
    import gpu_common as gc
    for bn, st in ((128, 3), (128, 2), (128, 4), (128, 6), (256, 2), (256, 4)):
        os.environ["CONZIC_GEMM_BN"], os.environ["CONZIC_GEMM_STAGES"] = str(bn), str(st)
        gc.drop_engines()
        eng = gc.engine("bf16", "tcgen05")
        for (M, N, K) in ((16384, 1536, 512), (16384, 512, 512), (16384, 2048, 512), (16384, 512, 2048),
                          (65536, 2048, 512), (65536, 512, 2048)):
            A = torch.randn(M, K, device="cuda")
            W = torch.randn(N, K, device="cuda") * 0.05
            ms = _time(lambda: eng.debug_linear(A, W, None, None, 0))
            # debug_linear converts operands each call; time the conversions alone and subtract
            emit("perf_gemm", dict(bn=bn, stages=st, M=M, N=N, K=K, ms_incl_convert=ms,
                                   tflops_incl_convert=2.0 * M * N * K / ms / 1e9))


def stage_persist(cg):
    """Persistent GEMM (CONZIC_GEMM_PERSIST=1, CONZIC_GEMM_CG=cg): correctness against fp64, then GEMM-only time
    from the library's per-launch CUDA events (operand conversion excluded)."""
    import torch
    import gpu_common as gc
    os.environ["CONZIC_GEMM_PERSIST"] = "1" if cg else "0"
    os.environ["CONZIC_GEMM_CG"] = str(cg or 1)
    eng = gc.engine("bf16", "tcgen05")
    for (M, N, K, act) in ((128, 256, 64, 0), (256, 256, 128, 0), (200, 512, 512, 1), (1000, 1536, 512, 0),
                           (333, 2048, 512, 1), (4096, 512, 2048, 0), (20000, 512, 512, 0), (37, 768, 768, 2),
                           (5000, 30524, 768, 0)):
        g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
        A = torch.randn(M, K, device="cuda", generator=g)
        W = torch.randn(N, K, device="cuda", generator=g) * 0.05
        bias = torch.randn(N, device="cuda", generator=g)
        resid = torch.randn(M, N, device="cuda", generator=g) if N % 4 == 0 else None
        out = eng.debug_linear(A, W, bias, resid, act)
        torch.cuda.synchronize()
        ref = ref_linear(A, W, bias, resid, act, True)
        err = float((out - ref).abs().max())
        emit("persist", dict(cg=cg, M=M, N=N, K=K, act=act, max_err=err, ref_max=float(ref.abs().max()),
                             ok=bool(err < 3e-4 * float(ref.abs().max()))))
        if N % 8:
            continue
        out = eng.debug_linear(A, W, bias, None, act | 16)  # bf16 activation output path
        torch.cuda.synchronize()
        ref = ref_linear(A, W, bias, None, act, True).bfloat16().float()
        err = float((out - ref).abs().max())
        emit("persist_bf16out", dict(cg=cg, M=M, N=N, K=K, act=act, max_err=err, ref_max=float(ref.abs().max()),
                                     ok=bool(err < 1.6e-2 * float(ref.abs().max()))))
    for (M, N, K) in ((18944, 1536, 512), (18944, 512, 512), (18944, 2048, 512), (18944, 512, 2048),
                      (75776, 1536, 512), (75776, 512, 512), (75776, 2048, 512), (75776, 512, 2048)):
        A = torch.randn(M, K, device="cuda")
        W = torch.randn(N, K, device="cuda") * 0.05
        bias = torch.randn(N, device="cuda")
        resid = torch.randn(M, N, device="cuda")
        for mode, (rs, act) in (("f32", (None, 0)), ("f32+resid", (resid, 0)), ("bf16+gelu", (None, 17)), ("bf16", (None, 16))):
            for _ in range(3):
                eng.debug_linear(A, W, bias, rs, act)
            eng.profile(True)
            for _ in range(10):
                eng.debug_linear(A, W, bias, rs, act)
            ms, work, n = eng.profile_read()["gemm"]
            eng.profile(False)
            emit("persist_perf", dict(cg=cg, mode=mode, M=M, N=N, K=K, ms=round(ms / n, 4),
                                      tflops=round(work / (ms / 1e3) / 1e12, 1)))


def stage_perf_step():
    import torch
    import gpu_common as gc
    from conzic_b200 import synth
    for prec in ("bf16", "bf16x3"):
        gc.drop_engines()
        eng = gc.engine(prec, "tcgen05")
        B, n, K = 64, 10, 200
        L = n + 5
        img = torch.randn(B, 512, device="cuda")
        for visited in (0, 5, 9):
            inp = torch.tensor([[101, 3746, 1997, 1037] + [103] * n + [102]] * B, device="cuda")
            for j in range(visited):
                inp[:, 4 + j] = 2000 + j * 13
            tm = synth.make_token_mask("cuda")
            pos = 4 + visited if visited < n else 4

            def run():
                eng.gibbs_step(inp, tm, img, pos, False, K, 0.1, 0.02, 2.0, 3 + visited, 0)
            ms = _time(run, iters=5, warm=2)
            emit("perf_step", dict(prec=prec, B=B, K=K, first_sweep_visited=visited, ms=ms))
        # full-length step in a later sweep: every other position holds a word
        inp = torch.tensor([[101, 3746, 1997, 1037] + [2000 + 7 * j for j in range(n)] + [102]] * B, device="cuda")
        tm = synth.make_token_mask("cuda")
        for ii in (0, 5, 9):
            def run():
                eng.gibbs_step(inp, tm, img, 4 + ii, ii == n - 1, K, 0.1, 0.02, 2.0, 3 + ii, n - 1 - ii)
            ms = _time(run, iters=5, warm=2)
            emit("perf_step", dict(prec=prec, B=B, K=K, later_sweep_ii=ii, ms=ms, launches=eng.launch_count()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage")
    ap.add_argument("--timeout", type=int, default=240)
    ap.add_argument("--stages", default=",".join(STAGES))
    a = ap.parse_args()
    if a.stage:
        fn = {"gemm": stage_gemm, "gemm_simt": lambda: stage_gemm("simt_debug"), "bert": stage_bert, "topk": stage_topk,
              "clip": stage_clip, "step": stage_step, "perf_gemm": stage_perf_gemm, "perf_step": stage_perf_step, "persist0": lambda: stage_persist(0),
              "persist1": lambda: stage_persist(1), "persist2": lambda: stage_persist(2)}[a.stage]
        fn()
        return
    for st in a.stages.split(","):
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st], timeout=a.timeout,
                               capture_output=True, text=True)
            tail = (r.stdout[-6000:] + "\n--- stderr ---\n" + r.stderr[-3000:])
            rc = r.returncode
        except subprocess.TimeoutExpired as e:
            tail, rc = f"TIMEOUT after {a.timeout}s\n" + ((e.stdout or b"")[-3000:].decode(errors="replace") if isinstance(e.stdout, bytes) else str(e.stdout)[-3000:]), -9
        print(f"===== stage {st}: rc={rc} {time.time()-t0:.1f}s =====\n{tail}", flush=True)
        emit("driver", dict(which=st, rc=rc, seconds=round(time.time() - t0, 1)))


if __name__ == "__main__":
    main()
