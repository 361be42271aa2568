"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the C ABI of libconzic.so;
the CPU oracle and the golden fixtures recorded from the unmodified reference are the checkers.

Tolerances (stated here; measured margins in profiles/r01h_parity.md and profiles/r02*_parity.md, from
tools/parity_report.py):
  * bf16x3 mode (3-pass split operands, the exact mode): BERT row logits within 1e-3 of the reference,
    CLIP cosine within 2e-5, softmax_K score within 2e-4; top-k ids identical wherever the reference's
    probabilities are non-zero and distinct; chosen token ids identical.
  * certified mode (the default: BERT bf16x3, CLIP bf16 + certified argmax with exact re-score): top-k ids,
    chosen token ids and the reported cosines identical to bf16x3's, hence to the reference's.
  * bf16 mode (everything bf16): logits within 0.04, cosine within 1.6e-3 (2x the measured maxima 1.8e-2 / 7.9e-4);
    chosen ids must agree whenever the reference's own top-2 margin exceeds what that cosine error can move.
"""
import random

import numpy as np
import pytest
import torch

import gpu_common as gc
from synthetic import synth

pytestmark = pytest.mark.gpu

FIXTURES = ["seq_b2_n4_k8", "shuffle_b3_n5_k16_multi", "senti_shuffle_neg_b2_n4_k8", "peaked_seq_b2_n4_k32",
            "random_b2_n3_k8", "senti_seq_b2_n4_k8", "seq_b1_n10_k200"]


def _ref_linear(A, W, bias, resid, act, bf16_round):
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


@pytest.mark.parametrize("prec", ["bf16", "bf16x3"])
@pytest.mark.parametrize("shape", [(128, 128, 64, 0), (200, 512, 512, 1), (1000, 1536, 512, 0), (333, 2048, 512, 1),
                                   (4096, 512, 2048, 0), (64, 30524, 768, 0), (18, 768, 768, 2), (1, 512, 512, 0)])
def test_tcgen05_linear_matches_fp64(prec, shape):
    """The GEMM both towers are made of, against an fp64 torch matmul of the same operands."""
    M, N, K, act = shape
    eng = gc.engine(prec, "tcgen05")
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    out = eng.debug_linear(A, W, bias, resid, act)
    ref = _ref_linear(A, W, bias, resid, act, prec == "bf16")
    tol = (3e-4 if prec == "bf16" else 3e-4) * float(ref.abs().max())
    assert float((out - ref).abs().max()) < tol


def test_tcgen05_matches_simt_debug_kernel():
    a = gc.engine("bf16", "tcgen05")
    b = gc.engine("bf16", "simt_debug")
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(300, 512, device="cuda", generator=g)
    W = torch.randn(640, 512, device="cuda", generator=g) * 0.05
    torch.testing.assert_close(a.debug_linear(A, W), b.debug_linear(A, W), rtol=0, atol=2e-4)


@pytest.mark.parametrize("mode", ["f32", "f32+resid", "bf16", "bf16+gelu"])
@pytest.mark.parametrize("shape", [(700, 1536, 512), (20000, 512, 512), (333, 2048, 512), (4096, 512, 2048),
                                   (75776, 512, 512)])
def test_persistent_pair_gemm_all_epilogues(mode, shape):
    """gemm_persist_kernel (CTA pairs, cta_group::2): every epilogue path against fp64 on bf16-rounded operands;
    M tails, A-resident (K=512) and streamed (K=2048) operands, whole-wave M."""
    M, N, K = shape
    eng = gc.engine("bf16", "tcgen05")
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) if mode == "f32+resid" else None
    act = {"f32": 0, "f32+resid": 0, "bf16": 16, "bf16+gelu": 17}[mode]
    out = eng.debug_linear(A, W, bias, resid, act)
    ref = _ref_linear(A, W, bias, resid, act & 15, True)
    if act & 16:  # bf16 output: one bf16 ulp (2^-8 relative) of slack; quick-gelu uses tanh.approx (~2^-11)
        assert float(((out - ref).abs() / ref.abs().clamp_min(1.0)).max()) < 2 ** -7
    else:
        assert float((out - ref).abs().max()) < 3e-4 * float(ref.abs().max())


@pytest.mark.parametrize("shape", [(4096 + 77, 512, 2048), (20000 + 33, 512, 512), (300, 512, 512)])
def test_wide_pair_gemm_with_fused_layernorm_inputs(shape):
    """gemm_wide_kernel (one 512-column unit per CTA pair, the kernel that also writes the LayerNorm of its rows on
    the CLIP tower's default path): fp32 + residual output against fp64, ragged M."""
    M, N, K = shape
    eng = gc.engine("bf16", "tcgen05")
    g = torch.Generator(device="cuda").manual_seed(M)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    out = eng.debug_linear(A, W, bias, resid, 0)
    ref = _ref_linear(A, W, bias, resid, 0, True)
    assert float((out - ref).abs().max()) < 3e-4 * float(ref.abs().max())


@pytest.mark.parametrize("shape", [(200, 512, 512, 1), (333, 2048, 512, 1), (1000, 512, 2048, 0)])
def test_certified_context_exact_tower_linear(shape):
    """A CERTIFIED context holds the CLIP linears twice: act | 32 routes debug_linear through the exact (bf16x3)
    copy's operand format and kernel, which must be as accurate as the bf16x3 context's."""
    M, N, K, act = shape
    eng = gc.engine("certified", "tcgen05")
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    out = eng.debug_linear(A, W, bias, resid, act | 32)
    ref = _ref_linear(A, W, bias, resid, act, False)
    assert float((out - ref).abs().max()) < 3e-4 * float(ref.abs().max())
    fast = eng.debug_linear(A, W, bias, resid, act)
    ref16 = _ref_linear(A, W, bias, resid, act, True)
    assert float((fast - ref16).abs().max()) < 3e-4 * float(ref16.abs().max())


@pytest.mark.parametrize("shape", [(2000, 512, 512, 0), (2100, 512, 2048, 0), (960, 2304, 768, 0), (5000, 2048, 512, 1),
                                   (300, 1536, 512, 0)])
def test_bf16x3_pair_and_gridded_kernels_are_bit_identical(shape):
    """The bf16x3 GEMM runs on the CTA-pair kernel or on the gridded 128 x 128 kernel depending on how many tiles the
    shape has (the certified re-score is a few thousand rows, the bf16x3 mode's full pass 100 k+): both accumulate the
    same k blocks in the same order, so their results must agree bit for bit -- which is what lets the certified mode
    reproduce the bf16x3 mode's scores exactly."""
    M, N, K, act = shape
    eng = gc.engine("bf16x3", "tcgen05")
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    pair = eng.debug_linear(A, W, bias, resid, act | 128)
    grid = eng.debug_linear(A, W, bias, resid, act | 64)
    auto = eng.debug_linear(A, W, bias, resid, act)
    assert torch.equal(pair, grid) and torch.equal(auto, pair)
    ref = _ref_linear(A, W, bias, resid, act, False)
    assert float((pair - ref).abs().max()) < 3e-4 * float(ref.abs().max())


@pytest.mark.parametrize("prec,tol", [("bf16x3", 1e-3), ("certified", 1e-3), ("bf16", 0.04)])
def test_bert_row_logits_vs_oracle(prec, tol):
    from oracle import conzic_oracle as orc
    sd = gc.weights("bert")
    inp = torch.tensor([[101, 3746, 1997, 1037, 103, 103, 103, 103, 102],
                        [101, 3746, 1997, 1037, 5000, 103, 7000, 2500, 102],
                        [101, 3746, 1997, 1037, 0, 103, 1012, 2500, 102]])
    with torch.no_grad():
        ref = orc.bert_mlm_head(sd, orc.bert_encoder(sd, inp)[:, 5])
    out = gc.engine(prec).bert_mlm_row(inp.cuda(), 5).cpu()
    assert float((out - ref).abs().max()) < tol


@pytest.mark.parametrize("prec,tol", [("bf16x3", 2e-4), ("bf16", 0.03), ("certified", 0.03)])
def test_clip_text_encode_vs_oracle(prec, tol):
    """Ragged lengths, EOS padding, T up to the 77-token cap."""
    from oracle import conzic_oracle as orc
    sd = gc.weights("clip")
    torch.manual_seed(1)
    for N, T in ((37, 11), (5, 77), (130, 6)):
        ids = torch.randint(300, 40000, (N, T))
        ids[:, 0] = synth.CLIP_BOS
        lens = torch.randint(2, T + 1, (N,))
        for i in range(N):
            ids[i, lens[i] - 1:] = synth.CLIP_EOS
        with torch.no_grad():
            ref = orc.clip_text_embeds(sd, ids)
        out = gc.engine(prec).clip_text_encode(ids.int().cuda()).cpu()
        assert float((out - ref).abs().max()) < tol


@pytest.mark.parametrize("prec,tol,kw", [("bf16x3", 3e-4, {}), ("bf16", 0.04, {}),
                                         ("bf16", 0.04, {"ln_standalone": True}), ("bf16", 0.04, {"pdl": False}),
                                         ("bf16", 0.04, {"wide_lsu": True}), ("bf16", 0.04, {"lsu_out": True})])
def test_clip_text_encode_with_nontrivial_layernorm(prec, tol, kw):
    """CLIP tower with perturbed LayerNorm gamma / beta (the synthetic checkpoint has gamma = 1, beta = 0) against
    the oracle: the default path (LayerNorm written by the O-proj / fc2 epilogues of the wide pair kernel and by the
    embedding kernel), the stand-alone LayerNorm kernels (ln_standalone) and launches without PDL."""
    from conzic_b200.engine import Engine
    from oracle import conzic_oracle as orc
    sd = {k: v.clone() for k, v in gc.weights("clip").items()}
    g = torch.Generator().manual_seed(11)
    for k in sd:
        if "layer_norm" in k and k.startswith("text_model"):
            if k.endswith(".weight"):
                sd[k] = 1.0 + 0.3 * torch.randn(sd[k].shape, generator=g)
            else:
                sd[k] = 0.2 * torch.randn(sd[k].shape, generator=g)
    eng = Engine(gc.weights("bert"), sd, device="cuda:0", precision=prec, **kw)
    torch.manual_seed(3)
    for N, T in ((300, 9), (41, 16)):
        ids = torch.randint(300, 40000, (N, T))
        ids[:, 0] = synth.CLIP_BOS
        lens = torch.randint(2, T + 1, (N,))
        for i in range(N):
            ids[i, lens[i] - 1:] = synth.CLIP_EOS
        with torch.no_grad():
            ref = orc.clip_text_embeds(sd, ids)
        out = eng.clip_text_encode(ids.int().cuda()).cpu()
        assert float((out - ref).abs().max()) < tol
    eng.close()


def test_gemm_epilogue_forms_agree():
    """The TMA epilogues against the per-lane ones they replace.  N = 512 GEMM: the reduce-add form (the L2 adds
    acc + bias to the residual stream) computes the same fp32 sums as the per-lane form -- one add per element, fp32
    addition commutes -- and LayerNorm statistics that differ only in summation order, so a rare bf16 rounding of a
    LayerNorm output is all that can differ.  bf16 outputs of the persistent kernel as TMA boxes: same bytes.  Checked
    on a 2-block tower (block 0: all rows through both fused-LayerNorm GEMMs; block 1: the compacted EOS rows) -- over
    12 blocks such roundings grow to the bf16 tower's own noise (~1e-3 on every element), which the oracle comparison
    above covers.  Ragged row counts exercise the clipped last tile."""
    from conzic_b200.engine import Engine
    sd = {k: v for k, v in gc.weights("clip").items()
          if ".layers." not in k or k.split(".layers.")[1].split(".")[0] in ("0", "1")}
    torch.manual_seed(5)
    cases = []
    for N, T in ((700, 9), (37, 12), (1, 5)):
        ids = torch.randint(300, 40000, (N, T))
        ids[:, 0] = synth.CLIP_BOS
        lens = torch.randint(2, T + 1, (N,))
        for i in range(N):
            ids[i, lens[i] - 1:] = synth.CLIP_EOS
        cases.append(ids.int().cuda())
    outs = {}
    for name, kw in (("tma", {}), ("wide_lsu", {"wide_lsu": True}), ("lsu_out", {"lsu_out": True})):
        eng = Engine(gc.weights("bert"), sd, device="cuda:0", precision="bf16", **kw)
        outs[name] = [eng.clip_text_encode(ids).cpu() for ids in cases]
        eng.close()
    for a, b in zip(outs["tma"], outs["lsu_out"]):
        assert torch.equal(a, b)
    for a, b in zip(outs["tma"], outs["wide_lsu"]):
        assert torch.isfinite(a).all()
        d = (a - b).abs()
        assert float(d.max()) < 2e-2                      # one flipped bf16 rounding moves an embedding by ~5e-3
        assert float((d.amax(dim=1) > 0).float().mean()) < 0.1 or a.shape[0] < 20   # ... and most rows not at all
        assert float(d.mean()) < 1e-4


@pytest.mark.parametrize("prec,tol", [("bf16x3", 1e-4), ("certified", 1e-4), ("bf16", 2e-2)])
def test_clip_image_encode_vs_oracle(prec, tol):
    """CLIP ViT-B/32 image tower on the engine's kernels (im2col + patch GEMM, pre-LN, 12 blocks with
    bidirectional 50-token attention, post-LN of the class token, projection) against the oracle's
    compute_image_representation (clip/clip.py:48-62)."""
    from conzic_b200.engine import Engine
    from oracle import conzic_oracle as orc
    sd = synth.make_clip_state_dict(0, vision=True)
    eng = Engine(gc.weights("bert"), sd, device="cuda:0", precision=prec)
    pix = torch.stack([synth.make_pixel_values(i) for i in range(5)])
    o = orc.Oracle(gc.weights("bert"), sd, synth.SynthBertTokenizer(), synth.SynthCLIPTokenizer(), full_logits=False)
    with torch.no_grad():
        ref = o.compute_image_representation(pix)
    out = eng.image_encode(pix.cuda()).cpu()
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) < tol * float(ref.abs().max())
    cos = torch.nn.functional.cosine_similarity(out, ref, dim=-1)
    assert float((1 - cos).max()) < (2e-4 if prec == "bf16" else 1e-7)
    eng.close()


def test_similarity_vs_oracle():
    from oracle import conzic_oracle as orc
    eng = gc.engine("bf16x3")
    torch.manual_seed(2)
    img, txt = torch.randn(3, 512), torch.randn(3 * 12, 512)
    rs, rr = orc.image_text_similarity(img, txt, torch.tensor(synth.LOGIT_SCALE))
    gs, gr = eng.image_text_similarity(img.cuda(), txt.cuda())
    torch.testing.assert_close(gs.cpu(), rs, rtol=0, atol=2e-6)
    torch.testing.assert_close(gr.cpu(), rr, rtol=0, atol=2e-7)


@pytest.mark.parametrize("K", [1, 8, 200, 512, 1000])
def test_topk_mask_exact_ids_and_tie_contract(K):
    """ids bit-exact where probabilities are non-zero; zero-probability ties come out in ascending id order
    starting from the lowest ids (the contract that replaces torch.topk's unspecified tie order)."""
    eng = gc.engine("bf16x3")
    V = synth.BERT_VOCAB
    torch.manual_seed(0)
    for case in ("smooth", "peaked", "sparse"):
        if case == "sparse":  # a few live tokens, everything else underflows to exactly 0 on any device
            logits = torch.where(torch.rand(3, V) < 0.002, torch.randn(3, V) * 0.3, torch.full((3, V), -30.0))
        else:
            logits = torch.randn(3, V) * (0.56 if case == "smooth" else 3.5)
        mask = synth.make_token_mask()
        probs = torch.softmax(logits / 0.1, dim=-1) * mask
        rp1, ri1 = probs.topk(K + 1, dim=-1)
        rp, ri = rp1[:, :K], ri1[:, :K]
        gp, gi = eng.topk_mask(logits.cuda(), mask.cuda(), 0.1, K)
        gp, gi = gp.cpu(), gi.cpu()
        nz = rp > 0
        # a rank is comparable id-for-id only if its probability is separated from both neighbours (rank K+1
        # included) by more than the few-ulp difference between the device's and the host's expf
        rel = (rp1[:, :-1] - rp1[:, 1:]) / rp1[:, :-1].clamp_min(1e-37)
        sep = rel > 1e-5
        distinct = sep.clone()
        distinct[:, 1:] &= sep[:, :-1]
        cmp = nz & distinct
        assert int(cmp.sum()) > 0.9 * int(nz.sum())
        assert torch.equal(gi[cmp], ri[cmp])
        torch.testing.assert_close(gp[nz], rp[nz], rtol=2e-5, atol=0)
        assert bool((gp[:, 1:] <= gp[:, :-1]).all())
        for r in range(3):
            z = gi[r][~nz[r]]
            if z.numel():
                zero_ids = torch.nonzero(probs[r] == 0).flatten()
                assert torch.equal(z, zero_ids[: z.numel()])


def test_build_clip_ids_matches_string_round_trip():
    """Device CSR assembly == tokenizer.batch_decode + CLIP tokenizer of the oracle (multi-token words,
    masked candidates that vanish, specials dropped)."""
    eng = gc.engine("bf16x3", multi=True)
    tok, ctok = synth.SynthBertTokenizer(), synth.SynthCLIPTokenizer(True)
    inp = torch.tensor([[101, 3746, 1997, 1037, 2003, 103, 103, 7003, 102],
                        [101, 3746, 1997, 1037, 0, 103, 1012, 2500, 102]])
    ids = torch.tensor([[2010, 5, 7010, 1012], [3000, 2999, 100, 4003]])
    mask = synth.make_token_mask()
    mask[0, 1012] = 1
    pos = 5
    idm = (ids * mask[0][ids]).long()
    cand = inp.unsqueeze(1).repeat(1, 4, 1)
    cand[:, :, pos] = idm
    texts = tok.batch_decode(cand.view(-1, inp.shape[1]), skip_special_tokens=True)
    ref = ctok(texts)["input_ids"]
    T = ref.shape[1] + 2
    got, glen, gidm = eng.build_clip_ids(inp.cuda(), pos, ids.cuda(), mask.cuda(), T)
    got, glen = got.cpu(), glen.cpu()
    assert torch.equal(gidm.cpu(), idm)
    assert torch.equal(got[:, : ref.shape[1]].long(), ref)
    assert bool((got[:, ref.shape[1]:] == synth.CLIP_EOS).all())
    ref_len = (ref == synth.CLIP_EOS).int().argmax(1) + 1
    assert torch.equal(glen.long(), ref_len)


@pytest.mark.parametrize("name", FIXTURES)
def test_gibbs_step_teacher_forced_bf16x3(name):
    """Every recorded step of the unmodified reference, fed its own `inp`: one conzic_gibbs_step must give the
    same logits / top-k ids / cosine / softmax / winner."""
    g = gc.load_golden(name)
    case = g["case"]
    eng = gc.engine("bf16x3", "tcgen05", case.get("peaked", False), case.get("multi", False))
    m = gc.replay_fixture(eng, g)
    assert m["dot_mismatch"] == 0
    assert m["logit_err"] < 1e-3
    assert m["topk_id_mismatch"] == 0
    assert m["clip_ref_err"] < 2e-5
    assert m["clip_score_err"] < 2e-4
    assert m["winner_mismatch"] == 0
    assert m["winner_checked"] > 0
    assert m["winner_cos_err"] < 2e-5


@pytest.mark.parametrize("kw", [{}, {"cert_zratio": (-1.0, -1.0)}, {"cert_dcos": 1.0}, {"cert_dcos": 0.05, "cert_fcap": 2},
                                {"cert_dcos": 0.02, "cert_fcap": 64}])
@pytest.mark.parametrize("name", FIXTURES)
def test_gibbs_step_teacher_forced_certified(name, kw):
    """The default precision, teacher-forced on every recorded step of the unmodified reference: exact logits and
    top-k ids (BERT runs in bf16x3), the reference's winner and its cosine (certified argmax).  Besides the default
    error bound: a bound so loose that every image takes the full exact re-encode (cert_dcos 1.0), a loose bound with
    a tiny survivor cap (round 2 and the overflow route), and a loose bound with a large cap (round 2 orders many
    survivors) -- every route must land on the same winners."""
    g = gc.load_golden(name)
    case = g["case"]
    eng = gc.engine("certified", "tcgen05", case.get("peaked", False), case.get("multi", False), **kw)
    before = eng.cert_stats()
    m = gc.replay_fixture(eng, g)
    st = {k: v - before[k] for k, v in eng.cert_stats().items()}
    assert m["dot_mismatch"] == 0
    assert m["logit_err"] < 1e-3
    assert m["topk_id_mismatch"] == 0
    assert m["winner_mismatch"] == 0
    assert m["winner_checked"] > 0
    assert m["winner_cos_err"] < 2e-5
    assert st["calls"] == m["steps"] and st["rescored_candidates"] >= 0
    if kw.get("cert_dcos") == 1.0 and case["K"] > 1:
        assert st["images_full"] > 0  # nothing can be ruled out with a bound that loose


@pytest.mark.parametrize("name", FIXTURES)
def test_gibbs_step_teacher_forced_bf16(name):
    g = gc.load_golden(name)
    case = g["case"]
    eng = gc.engine("bf16", "tcgen05", case.get("peaked", False), case.get("multi", False))
    m = gc.replay_fixture(eng, g)
    assert m["logit_err"] < 0.04
    assert m["cos_checked"] > 0 and m["clip_ref_err"] < 1.6e-3
    # winners are checked image by image wherever the top-k id set matches the reference's; a flipped winner is only
    # acceptable on a near tie of the fused score: beta * (softmax change a 1.6e-3 cosine error can cause) < 0.7
    assert m["winner_mismatch"] == 0 or m["min_margin_at_mismatch"] < 0.7


def _models(case):
    from conzic_b200.clip.clip import CLIP
    from conzic_b200.models import BertMLM
    bert = BertMLM(gc.weights("bert", case.get("peaked", False)))
    clip = CLIP(state_dict=gc.weights("clip"), tokenizer=synth.SynthCLIPTokenizer(case.get("multi", False)),
                processor=synth.SynthProcessor())
    return bert, clip.to("cuda:0")


@pytest.mark.parametrize("name", ["seq_b2_n4_k8", "shuffle_b3_n5_k16_multi", "random_b2_n3_k8", "senti_seq_b2_n4_k8",
                                  "senti_shuffle_neg_b2_n4_k8", "peaked_seq_b2_n4_k32", "span_b2_n5_k8",
                                  "pos_seq_b2_n5_k16"])
@pytest.mark.parametrize("prec", ["certified", "bf16x3"])
def test_free_running_call_matches_reference(name, prec, monkeypatch):
    """generate_caption / control_generate_caption through the drop-in API under set_seed(42): same captions per
    sweep, same best list, same CLIP scores as the unmodified reference returned -- in the default (certified)
    precision and in bf16x3."""
    import logging
    from conzic_b200 import control_gen_utils, gen_utils, runtime
    from conzic_b200.utils import set_seed
    monkeypatch.setenv("CONZIC_PRECISION", prec)
    runtime.clear()
    control_gen_utils.set_pos_tagger(synth.synth_pos_tagger)  # POS control: the tagger is a plug-in
    g = gc.load_golden(name)
    case = g["case"]
    bert, clip = _models(case)
    B, n, K = case["B"], case["n"], case["K"]
    pix = torch.stack([synth.make_pixel_values(i) for i in range(B)])
    token_mask = synth.make_token_mask("cuda")
    logger = logging.getLogger("test")
    names = [f"img{i}.jpg" for i in range(B)]
    kw = dict(prompt=synth.SYNTH_PROMPT, batch_size=B, max_len=n, top_k=K, temperature=0.1, max_iter=case["iters"],
              alpha=0.02, beta=2.0, generate_order=case["order"])
    set_seed(42)
    if case.get("gamma") is None:
        texts, scores = gen_utils.generate_caption(names, bert, clip, synth.SynthBertTokenizer(), pix, token_mask,
                                                   logger, **kw)
    else:
        texts, scores = control_gen_utils.control_generate_caption(
            names, bert, clip, synth.SynthBertTokenizer(), pix, token_mask, logger, gamma=case["gamma"],
            ctl_type=case.get("ctl", "sentiment"), style_type=case.get("style", "positive"),
            pos_type=synth.SYNTH_POS_TEMPLATE, sentiment_table=synth.make_sentiment_table(), **kw)
    runtime.clear()
    assert texts == g["texts"]
    for a, b in zip(scores, g["scores"]):
        np.testing.assert_allclose(a, b, rtol=0, atol=2e-5)


@pytest.mark.parametrize("name", ["seq_b2_n4_k8", "shuffle_b3_n5_k16_multi", "senti_shuffle_neg_b2_n4_k8", "span_b2_n5_k8"])
@pytest.mark.parametrize("prec", ["certified", "bf16x3"])
def test_host_string_path_matches_reference(name, prec, monkeypatch):
    """CONZIC_STRING_PATH=1: candidate ids -> host -> batch_decode -> CLIP tokenizer -> device, every arithmetic
    piece still a libconzic kernel (conzic_score_select with the candidates' CLIP ids for the certified re-score)
    gives the reference's captions and scores too."""
    import logging
    from conzic_b200 import control_gen_utils, gen_utils, runtime
    from conzic_b200.utils import set_seed
    monkeypatch.setenv("CONZIC_PRECISION", prec)
    monkeypatch.setenv("CONZIC_STRING_PATH", "1")
    runtime.clear()
    g = gc.load_golden(name)
    case = g["case"]
    bert, clip = _models(case)
    B, n, K = case["B"], case["n"], case["K"]
    pix = torch.stack([synth.make_pixel_values(i) for i in range(B)])
    kw = dict(prompt=synth.SYNTH_PROMPT, batch_size=B, max_len=n, top_k=K, temperature=0.1, max_iter=case["iters"],
              alpha=0.02, beta=2.0, generate_order=case["order"])
    names = [f"img{i}.jpg" for i in range(B)]
    set_seed(42)
    if case.get("gamma") is None:
        texts, scores = gen_utils.generate_caption(names, bert, clip, synth.SynthBertTokenizer(), pix,
                                                   synth.make_token_mask("cuda"), logging.getLogger("test"), **kw)
    else:
        texts, scores = control_gen_utils.control_generate_caption(
            names, bert, clip, synth.SynthBertTokenizer(), pix, synth.make_token_mask("cuda"), logging.getLogger("test"),
            gamma=case["gamma"], ctl_type="sentiment", style_type=case["style"],
            sentiment_table=synth.make_sentiment_table(), **kw)
    runtime.clear()
    assert texts == g["texts"]
    for a, b in zip(scores, g["scores"]):
        np.testing.assert_allclose(a, b, rtol=0, atol=2e-5)


@pytest.mark.parametrize("name", ["pieces_seq_b2_n5_k16", "pieces_shuffle_b3_n6_k16_multi"])
@pytest.mark.parametrize("prec", ["certified", "bf16x3"])
def test_duck_typed_piece_vocabulary_matches_reference(name, prec, monkeypatch):
    """A duck-typed tokenizer pair whose BERT vocabulary has '##' word pieces (not the Hugging Face classes the
    device text pipeline restates): the engine detects that and takes the reference's string round trip every step
    (gen_utils.py:75); pieces among the candidates and, once one wins, inside the caption.  Must reproduce the
    unmodified reference's captions and scores."""
    import logging
    from conzic_b200 import gen_utils, runtime
    from conzic_b200.utils import set_seed
    monkeypatch.setenv("CONZIC_PRECISION", prec)
    runtime.clear()
    g = gc.load_golden(name)
    case = g["case"]
    from conzic_b200.clip.clip import CLIP
    from conzic_b200.models import BertMLM
    bert = BertMLM(gc.weights("bert"))
    clip = CLIP(state_dict=gc.weights("clip"), tokenizer=synth.PieceCLIPTokenizer(case.get("multi", False)),
                processor=synth.SynthProcessor()).to("cuda:0")
    B, n, K = case["B"], case["n"], case["K"]
    pix = torch.stack([synth.make_pixel_values(i) for i in range(B)])
    set_seed(42)
    bert_tok = synth.PieceBertTokenizer()  # engines are dropped when their tokenizer is collected: keep it alive
    texts, scores = gen_utils.generate_caption(
        [f"img{i}.jpg" for i in range(B)], bert, clip, bert_tok, pix, synth.make_token_mask("cuda"),
        logging.getLogger("test"), prompt=synth.SYNTH_PROMPT, batch_size=B, max_len=n, top_k=K, temperature=0.1,
        max_iter=case["iters"], alpha=0.02, beta=2.0, generate_order=case["order"])
    eng = runtime.any_engine()
    assert eng.needs_strings and not eng.has_text_vocab
    runtime.clear()
    if "shuffle" in name:  # in this fixture a piece wins a slot, so later steps see a merged word inside the caption
        assert any("p" in w[1:] for t in g["texts"] for c in t for w in c.split()), "fixture holds no merged word"
    assert texts == g["texts"]
    for a, b in zip(scores, g["scores"]):
        np.testing.assert_allclose(a, b, rtol=0, atol=2e-5)


@pytest.mark.parametrize("B,n,K", [(1, 1, 1), (2, 1, 1024), (1, 3, 5), (3, 2, 64)])
def test_extreme_shapes_step_vs_oracle(B, n, K):
    """Smallest and largest shapes the ABI accepts: one image / one candidate / one-word sentences (the only
    position is also the last, so '.' is allowed), and candidate_k at the 1024 cap.  Every position of one sweep
    is compared with the oracle in bf16x3 mode: top-k ids, cosines, winners."""
    from oracle import conzic_oracle as orc
    eng = gc.engine("bf16x3")
    o = orc.Oracle(gc.weights("bert"), gc.weights("clip"), synth.SynthBertTokenizer(), synth.SynthCLIPTokenizer(),
                   full_logits=False)
    o.trace = []
    tok = synth.SynthBertTokenizer()
    inp_ref = torch.tensor([tok.encode(synth.SYNTH_PROMPT + "[MASK]" * n)] * B)
    inp_dev = inp_ref.clone().cuda()
    img = torch.nn.functional.normalize(torch.randn(B, 512, generator=torch.Generator().manual_seed(21)), dim=-1)
    tm_ref, tm_dev = synth.make_token_mask(), synth.make_token_mask("cuda")
    for ii in range(n):
        pos = 4 + ii
        with torch.no_grad():
            o.step(inp_ref, img, tm_ref, pos, ii, n, K, 0.1, 0.02, 2.0)
        t = o.trace[-1]
        _, _, tr = eng.gibbs_step(inp_dev, tm_dev, img.cuda(), pos, ii == n - 1, K, 0.1, 0.02, 2.0, 3 + ii, 0, trace=True)
        torch.cuda.synchronize()
        p = t["probs"]
        rel = (p[:, :-1] - p[:, 1:]) / p[:, :-1].clamp_min(1e-30)
        ok = torch.ones_like(p, dtype=torch.bool)
        ok[:, :-1] &= rel > 1e-3
        ok[:, 1:] &= rel > 1e-3
        ok &= p > 0
        assert torch.equal(tr["idxs"].cpu()[ok], t["idxs"][ok])
        if torch.equal(tr["idxs"].cpu(), t["idxs"]):
            assert float((tr["clip_ref"].cpu() - t["clip_ref"]).abs().max()) < 2e-5
            assert torch.equal(inp_dev.cpu(), inp_ref)
        else:  # a near tie at the top-k boundary reordered two candidates: re-align and carry on
            inp_dev.copy_(inp_ref)
        assert float(tm_dev[0, synth.DOT_ID]) == float(tm_ref[0, synth.DOT_ID])


def test_cabi_rejects_bad_arguments():
    """Error convention of include/conzic.h: a negative return code and a message in conzic_last_error(), surfaced
    as RuntimeError by the binding; nothing is launched, the context stays usable."""
    eng = gc.engine("bf16x3")
    B, n = 2, 3
    tok = synth.SynthBertTokenizer()
    inp = torch.tensor([tok.encode(synth.SYNTH_PROMPT + "[MASK]" * n)] * B).cuda()
    tm = synth.make_token_mask("cuda")
    img = torch.randn(B, 512, device="cuda")
    with pytest.raises(RuntimeError, match="bad B / K / pos"):
        eng.gibbs_step(inp, tm, img, 4, False, 1025, 0.1, 0.02, 2.0, 3, 0)       # K over the 1024 cap
    with pytest.raises(RuntimeError, match="bad B / K / pos"):
        eng.gibbs_step(inp, tm, img, inp.shape[1] - 1, False, 8, 0.1, 0.02, 2.0, 3, 0)  # pos on [SEP]
    with pytest.raises(RuntimeError, match="bad B / K / pos"):
        eng.gibbs_step(inp, tm, img, 0, False, 8, 0.1, 0.02, 2.0, 0, 0)          # pos on [CLS]
    ws = torch.empty(1024, dtype=torch.uint8, device="cuda")
    out = torch.empty((B, eng.ldl), dtype=torch.float32, device="cuda")
    rc = eng.lib.conzic_bert_mlm_row(eng.ctx, inp.data_ptr(), B, inp.shape[1], 4, out.data_ptr(), eng.ldl, ws.data_ptr(),
                                     ws.numel(), None)
    from conzic_b200 import _lib
    assert rc < 0 and "workspace too small" in _lib.last_error()
    with pytest.raises(RuntimeError, match=r"T must be in \[1, 77\]"):
        eng.clip_text_encode(torch.full((2, 78), synth.CLIP_EOS, dtype=torch.int32))
    # still healthy afterwards
    _, _, tr = eng.gibbs_step(inp, tm, img, 4, False, 8, 0.1, 0.02, 2.0, 3, 0, trace=True)
    torch.cuda.synchronize()
    assert bool((inp[:, 4] >= 1996).all())


@pytest.mark.parametrize("vocab", ["letters", "rich"])
def test_device_text_kernel_matches_hf_tokenizers(vocab, tmp_path):
    """text_tokenize_kernel (csrc/text_ops.cu around text_pipeline.cuh) against the real transformers classes at
    scale: 48 images x 128 candidates, words / '##' pieces / punctuation / specials / masked candidates; every
    candidate caption's CLIP ids must equal CLIPTokenizer(BertTokenizer.batch_decode(ids, skip_special_tokens=True))
    (gen_utils.py:75 + clip/clip.py:71-72), BOS / EOS framing, EOS padding and lengths included."""
    from conzic_b200 import runtime, tokens
    from conzic_b200.engine import Engine
    if vocab == "letters":
        bert_tok, clip_tok = synth.make_hf_tokenizers(str(tmp_path))
        V, lo_dense, n_dense = synth.BERT_VOCAB, 1996, synth.BERT_VOCAB - 1996
    else:
        bert_tok, clip_tok = synth.make_hf_tokenizers_rich(str(tmp_path))
        V, lo_dense, n_dense = 3000, 1996, 116
    sd = runtime._dummy_bert_sd()
    sd["bert.embeddings.word_embeddings.weight"] = torch.zeros(V, 64)
    sd["cls.predictions.bias"] = torch.zeros(V)
    eng = Engine(sd, gc.weights("clip"), device="cuda:0", precision="bf16")
    special = list(synth.SPECIAL_IDS)
    off, tok, pieces = tokens.build_bert2clip(bert_tok, clip_tok, V, special)
    eng.set_bert2clip(off, tok)
    eng.set_text_vocab(tokens.build_text_vocab(bert_tok, clip_tok, V, special, off, tok))
    g = torch.Generator().manual_seed(7)
    B, L, K, pos = 48, 14, 128, 6
    inp = torch.randint(1996, V, (B, L), generator=g)
    dense = torch.rand((B, L), generator=g) < 0.6
    inp[dense] = torch.randint(lo_dense, lo_dense + n_dense, (int(dense.sum()),), generator=g)
    inp[torch.rand((B, L), generator=g) < 0.06] = synth.DOT_ID
    inp[torch.rand((B, L), generator=g) < 0.04] = synth.PAD_ID
    inp[:, 0], inp[:, -1] = synth.CLS_ID, synth.SEP_ID
    inp[:, pos] = synth.MASK_ID
    inp[0, 1:pos] = synth.PAD_ID  # the candidate becomes the first kept token: a piece candidate stays verbatim
    ids = torch.randint(lo_dense, lo_dense + n_dense, (B, K), generator=g)
    ids[:, ::7] = torch.randint(1996, V, (B, len(range(0, K, 7))), generator=g)
    ids[:, 3] = synth.DOT_ID
    mask = torch.ones(1, V)
    mask[0, ids[1, 5]] = 0  # a masked candidate: the word vanishes from the caption (gen_utils.py:72)
    idm = (ids * mask[0][ids]).long()
    cand = inp.unsqueeze(1).repeat(1, K, 1)
    cand[:, :, pos] = idm
    texts = bert_tok.batch_decode(cand.view(-1, L), skip_special_tokens=True)
    ref = clip_tok(texts, padding="max_length", max_length=77, truncation=True, return_tensors="pt")["input_ids"]
    got, glen, gidm = eng.build_clip_ids(inp.cuda(), pos, ids.cuda(), mask.cuda(), 77)
    torch.cuda.synchronize()
    assert torch.equal(gidm.cpu(), idm)
    # rows are compared up to the first EOS id (where the tower pools; with the letter vocabulary '#' is an unknown
    # symbol whose id is the EOS id, like in the released CLIP vocabulary), the rest must be EOS padding
    ref_len = (ref == synth.CLIP_EOS).int().argmax(1) + 1
    assert torch.equal(glen.cpu().long(), ref_len)
    live = torch.arange(77)[None, :] < ref_len[:, None]
    want = torch.where(live, ref, torch.full_like(ref, synth.CLIP_EOS))
    bad = (got.cpu().long() != want).any(dim=1).nonzero().flatten().tolist()
    assert not bad, (len(bad), texts[bad[0]], got[bad[0]].tolist()[:24], want[bad[0]].tolist()[:24])
    assert any("##" in t for t in texts) or vocab == "rich"
    eng.close()


def _hf_generate(prec, string_path, tmp_path, monkeypatch, B, n, K, iters, order, mode="caption", prompt=None):
    import logging
    from conzic_b200 import control_gen_utils, gen_utils, runtime
    from conzic_b200.clip.clip import CLIP
    from conzic_b200.models import BertMLM
    from conzic_b200.utils import set_seed
    monkeypatch.setenv("CONZIC_PRECISION", prec)
    if string_path:
        monkeypatch.setenv("CONZIC_STRING_PATH", "1")
    else:
        monkeypatch.delenv("CONZIC_STRING_PATH", raising=False)
    runtime.clear()
    bert_tok, clip_tok = synth.make_hf_tokenizers(str(tmp_path))
    bert = BertMLM(gc.weights("bert"))
    clip = CLIP(state_dict=gc.weights("clip"), tokenizer=clip_tok, processor=synth.SynthProcessor()).to("cuda:0")
    pix = torch.stack([synth.make_pixel_values(i) for i in range(B)])
    lines = []
    logger = logging.getLogger(f"hf-{prec}-{string_path}-{mode}-{order}")
    logger.setLevel(logging.INFO)
    logger.propagate = False
    h = logging.Handler()
    h.emit = lambda rec: lines.append(rec.getMessage())
    logger.addHandler(h)
    kw = dict(prompt=prompt or synth.hf_prompt(), batch_size=B, max_len=n, top_k=K, temperature=0.1, max_iter=iters,
              alpha=0.02, beta=2.0)
    names = [f"img{i}.jpg" for i in range(B)]
    set_seed(42)
    if mode == "sentiment":
        out = control_gen_utils.control_generate_caption(
            names, bert, clip, bert_tok, pix, synth.make_token_mask("cuda"), logger, gamma=5.0, ctl_type="sentiment",
            style_type="positive", generate_order=order, sentiment_table=synth.make_sentiment_table(), **kw)
    else:
        out = gen_utils.generate_caption(names, bert, clip, bert_tok, pix, synth.make_token_mask("cuda"), logger,
                                         generate_order=order, **kw)
    eng = runtime.any_engine()
    info = dict(has_text=eng.has_text_vocab, W=eng.max_tok_per_word, launches=eng.launch_count())
    runtime.clear()
    return out, [ln for ln in lines if ln.startswith("iter ")], info, clip_tok


@pytest.mark.parametrize("prec", ["certified", "bf16x3"])
def test_real_hf_tokenizer_classes_match_reference(prec, monkeypatch, tmp_path):
    """generate_caption with the REAL transformers BertTokenizer / CLIPTokenizer classes (generated vocabulary files:
    WordPiece decode with '##' joining and clean-up, byte-level BPE with merges, several CLIP tokens per word) against
    the fixture the unmodified reference produced with the same tokenizers.  The whole step runs on the device: the
    candidates' CLIP ids come from the device text pipeline (conzic_set_text_vocab), no string leaves the GPU."""
    g = gc.load_golden("hf_shuffle_b3_n5_k24")
    case = g["case"]
    (texts, scores), _, info, _ = _hf_generate(prec, False, tmp_path, monkeypatch, case["B"], case["n"], case["K"],
                                               case["iters"], case["order"])
    assert info["has_text"] and info["W"] > 3
    assert texts == g["texts"]
    for a, b in zip(scores, g["scores"]):
        np.testing.assert_allclose(a, b, rtol=0, atol=2e-5)


def test_long_token_heavy_captions_device_text_equals_string_path(monkeypatch, tmp_path):
    """Real tokenizer classes, 3-5 CLIP tokens per word, 22-word sentences: captions reach the 77-token truncation
    and the longest prefix + longest suffix exceed one attention tile (the step then shortens the shared prefix).
    The device text pipeline and the reference's string round trip must agree exactly."""
    B, n, K = 2, 22, 8
    out = {}
    for path in ("device", "strings"):
        (t, s), _, info, clip_tok = _hf_generate("bf16x3", path == "strings", tmp_path, monkeypatch, B, n, K, 2, "sequential")
        out[path] = (t, s)
    (ta, sa), (tb, sb) = out["device"], out["strings"]
    assert ta == tb
    for a, b in zip(sa, sb):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-6)
    n_tok = max(len(clip_tok(c, add_special_tokens=False)["input_ids"]) for c in ta[-2])
    assert n_tok > 75, "captions are not token heavy enough to exercise truncation / the tile guard"


@pytest.mark.parametrize("mode,order", [("sentiment", "shuffle"), ("caption", "span"), ("caption", "random"),
                                        ("caption", "sequential")])
def test_piece_vocabulary_device_text_equals_string_path(mode, order, monkeypatch, tmp_path):
    """Piece vocabulary (real tokenizer classes), modes the reference fixtures do not cover: the device text pipeline
    and the all-strings step are two routes to the same numbers -- captions, CLIP scores and the per-sweep log lines
    must agree exactly, in the default (certified) precision; the device route must not be the slow one."""
    B, n, K = 3, 6, 16
    res = {}
    for path in ("device", "strings"):
        out, lines, info, _ = _hf_generate("certified", path == "strings", tmp_path, monkeypatch, B, n, K, 2, order, mode)
        res[path] = (out, lines)
    (ta, sa), la = res["device"]
    (tb, sb), lb = res["strings"]
    assert ta == tb and la == lb and len(la) > 0
    for a, b in zip(sa, sb):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-6)


def test_long_sentence_free_running_matches_oracle():
    """sentence_len 25 (BASELINE config 5's longest): prefixes up to 29 tokens and candidate suffixes up to 27 rows
    -> the 64-key attention tiles and the multi-tile query path; multi-token words.  One sweep of a free-running
    call (bf16x3) against the CPU oracle on the same seeded inputs: identical token ids, scores within 2e-5."""
    import logging
    from conzic_b200 import gen_utils, runtime
    from conzic_b200.clip.clip import CLIP
    from conzic_b200.models import BertMLM
    from oracle import conzic_oracle as orc
    import os
    os.environ["CONZIC_PRECISION"] = "bf16x3"
    try:
        runtime.clear()
        B, n, K = 2, 25, 24
        bert_sd, clip_sd = gc.weights("bert"), synth.make_clip_state_dict(0, vision=True)
        o = orc.Oracle(bert_sd, clip_sd, synth.SynthBertTokenizer(), synth.SynthCLIPTokenizer(True), full_logits=False)
        pix = torch.stack([synth.make_pixel_values(i) for i in range(B)])
        with torch.no_grad():
            ref_texts, ref_scores = o.generate(pix, synth.make_token_mask(), synth.SYNTH_PROMPT, order="sequential",
                                               max_len=n, top_k=K, max_iters=1)
        bert = BertMLM(bert_sd)
        clip = CLIP(state_dict=clip_sd, tokenizer=synth.SynthCLIPTokenizer(True), processor=synth.SynthProcessor()).to("cuda:0")
        texts, scores = gen_utils.generate_caption([f"img{i}.jpg" for i in range(B)], bert, clip, synth.SynthBertTokenizer(),
                                                   pix, synth.make_token_mask("cuda"), logging.getLogger("test"),
                                                   prompt=synth.SYNTH_PROMPT, batch_size=B, max_len=n, top_k=K,
                                                   temperature=0.1, max_iter=1, alpha=0.02, beta=2.0,
                                                   generate_order="sequential")
        assert texts == ref_texts
        for a, b in zip(scores, ref_scores):
            np.testing.assert_allclose(a, b, rtol=0, atol=2e-5)
    finally:
        os.environ.pop("CONZIC_PRECISION", None)
        runtime.clear()


def test_prefix_sharing_equals_dense_encode():
    """Size-independent property: encoding candidates as shared prefix + per-candidate suffix gives the same
    cosine as encoding every full caption densely (causal tower => prefix states do not depend on the suffix)."""
    eng = gc.engine("bf16x3")
    B, n, K = 4, 6, 32
    L = n + 5
    torch.manual_seed(5)
    inp = torch.tensor([[101, 3746, 1997, 1037] + [2000 + 11 * j for j in range(n)] + [102]] * B).cuda()
    inp[1, 6] = 0
    tm = synth.make_token_mask("cuda")
    img = torch.randn(B, 512, device="cuda")
    pos = 7
    inp0 = inp.clone()
    _, _, tr = eng.gibbs_step(inp, tm, img, pos, False, K, 0.1, 0.02, 2.0, 3 + 3, n - 4, trace=True)
    masked = inp0.clone()
    masked[:, pos] = synth.MASK_ID
    cids, clen, idm = eng.build_clip_ids(masked, pos, tr["idxs"], tm, 20)
    emb = eng.clip_text_encode(cids)
    score, ref = eng.image_text_similarity(img, emb)
    torch.testing.assert_close(ref, tr["clip_ref"], rtol=0, atol=3e-6)
    torch.testing.assert_close(score, tr["clip_score"], rtol=0, atol=3e-5)


@pytest.mark.parametrize("prec,tol", [("bf16x3", 1e-5), ("bf16", 6e-3)])
def test_full_size_prefix_sharing_equals_dense_encode(prec, tol):
    """BASELINE config 2 sizes (B=64, K=200, len=10, middle position): the cosines of the shared-prefix step equal
    those of encoding all 12 800 candidate captions densely, token by token, through conzic_clip_text_encode."""
    eng = gc.engine(prec)
    B, n, K, ii = 64, 10, 200, 4
    pos = 4 + ii
    g = torch.Generator().manual_seed(17)
    words = torch.randint(2000, 30000, (B, n), generator=g)
    inp = torch.cat([torch.tensor([[101, 3746, 1997, 1037]] * B), words, torch.tensor([[102]] * B)], dim=1).cuda()
    inp[3, 6] = 0   # a dropped word in a prefix
    inp[5, 11] = 0  # and in a tail
    img = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1).cuda()
    tm = synth.make_token_mask("cuda")
    inp0 = inp.clone()
    _, _, tr = eng.gibbs_step(inp, tm, img, pos, False, K, 0.1, 0.02, 2.0, 3 + ii, n - 1 - ii, trace=True)
    masked = inp0.clone()
    masked[:, pos] = synth.MASK_ID
    cids, clen, idm = eng.build_clip_ids(masked, pos, tr["idxs"], tm, n + 6)
    assert int(clen.max()) <= n + 5 and int(clen.min()) >= n + 3
    emb = eng.clip_text_encode(cids)
    score, ref = eng.image_text_similarity(img, emb)
    assert float((ref - tr["clip_ref"]).abs().max()) < tol
    if prec == "bf16x3":
        best_dense = (0.02 * tr["probs"] + 2.0 * score).argmax(dim=1)
        assert torch.equal(best_dense, tr["best"])


def test_full_size_sentiment_step_properties():
    """BASELINE config 4 sizes (B=64 per GPU, sentence_len 12, K=200, sentiment gamma=5): legal winners, the control
    term follows the table (softmax over K of the caption's table sum), repeats penalise duplicates, and the step
    is deterministic."""
    eng = gc.engine("bf16x3")
    B, n, K, ii = 64, 12, 200, 5
    pos = 4 + ii
    g = torch.Generator().manual_seed(23)
    words = torch.randint(2000, 30000, (B, n), generator=g)
    inp0 = torch.cat([torch.tensor([[101, 3746, 1997, 1037]] * B), words, torch.tensor([[102]] * B)], dim=1).cuda()
    img = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1).cuda()
    table = synth.make_sentiment_table().cuda()
    runs = []
    for _ in range(2):
        inp = inp0.clone()
        tm = synth.make_token_mask("cuda")
        cr, se, tr = eng.gibbs_step(inp, tm, img, pos, False, K, 0.1, 0.02, 2.0, 3 + ii, n - 1 - ii, gamma=5.0,
                                    senti_table=table, trace=True)
        torch.cuda.synchronize()
        runs.append((inp.clone(), cr.clone(), se.clone(), tr))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1]) and torch.equal(runs[0][2], runs[1][2])
    inp, cr, se, tr = runs[0]
    assert bool((inp[:, pos] >= 1996).all())
    # rebuild the fused score from the traced pieces (control_gen_utils.py:53-59)
    cand = inp0.unsqueeze(1).repeat(1, K, 1)
    idm = (tr["idxs"] * synth.make_token_mask("cuda")[0][tr["idxs"]]).long()
    cand[:, :, pos] = idm
    special = torch.tensor(synth.SPECIAL_IDS, device="cuda")
    vis = ~torch.isin(cand, special)
    senti_raw = (table[cand] * vis).sum(-1)
    repeats = (idm[:, :, None] == cand).float().sum(2) - 1
    final = 0.02 * tr["probs"] + 2.0 * tr["clip_score"] + 5.0 * torch.softmax(senti_raw, dim=1) + 0.1 * (1 - torch.exp(repeats))
    torch.testing.assert_close(final, tr["final"], rtol=0, atol=2e-5)
    best = tr["best"]
    assert torch.equal(idm.gather(1, best.view(-1, 1)).squeeze(1), inp[:, pos])
    torch.testing.assert_close(se, senti_raw.gather(1, best.view(-1, 1)).squeeze(1), rtol=0, atol=1e-5)


def test_full_size_step_properties():
    """BASELINE config 2 sizes (B=64, K=200, len=10): winners are legal (unmasked) ids, scores are cosines, the step
    is deterministic; the certified precision gives bf16x3's winners and bit-identical reported cosines; the all-bf16
    precision picks the same winner on all but near ties."""
    B, n, K = 64, 10, 200
    img = torch.nn.functional.normalize(torch.randn(B, 512, generator=torch.Generator().manual_seed(9)), dim=-1).cuda()
    base = torch.tensor([[101, 3746, 1997, 1037] + [2000 + 7 * j for j in range(n)] + [102]] * B).cuda()
    outs = {}
    for prec in ("bf16x3", "certified", "bf16"):
        eng = gc.engine(prec)
        runs = []
        for _ in range(2):
            inp = base.clone()
            tm = synth.make_token_mask("cuda")
            cr, _, tr = eng.gibbs_step(inp, tm, img, 9, False, K, 0.1, 0.02, 2.0, 8, 4, trace=True)
            runs.append((inp.cpu(), cr.cpu(), tr["final"].cpu(), tr["idxs"].cpu()))
        assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
        w = runs[0][0][:, 9]
        assert bool((w >= 1996).all()) and bool((runs[0][1].abs() <= 1.0).all())
        outs[prec] = runs[0]
    assert torch.equal(outs["certified"][3], outs["bf16x3"][3]), "top-k ids differ (BERT must run in bf16x3)"
    assert torch.equal(outs["certified"][0], outs["bf16x3"][0]), "certified argmax picked another winner"
    assert torch.equal(outs["certified"][1], outs["bf16x3"][1]), "reported cosine is not the exact tower's"
    st = gc.engine("certified").cert_stats()
    assert st["rescored_candidates"] >= st["images"] > 0
    same = outs["bf16"][0][:, 9] == outs["bf16x3"][0][:, 9]
    top2 = outs["bf16x3"][2].topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    assert bool(same[margin > 0.7].all()), "bf16 flipped a winner whose fp32 margin was large"
    gc.drop_engines()


def _run_generate(prec, B, n, K, sweeps, order, monkeypatch):
    import logging
    from conzic_b200 import gen_utils, runtime
    from conzic_b200.clip.clip import CLIP
    from conzic_b200.models import BertMLM
    from conzic_b200.utils import set_seed
    monkeypatch.setenv("CONZIC_PRECISION", prec)
    runtime.clear()
    bert = BertMLM(gc.weights("bert"))
    clip = CLIP(state_dict=synth.make_clip_state_dict(0, vision=True), tokenizer=synth.SynthCLIPTokenizer(),
                processor=synth.SynthProcessor()).to("cuda")  # index-less device, like the reference's scripts
    pix = torch.stack([synth.make_pixel_values(i) for i in range(B)])
    set_seed(42)
    bert_tok = synth.SynthBertTokenizer()  # engines are dropped when their tokenizer is collected: keep it alive
    out = gen_utils.generate_caption([f"img{i}.jpg" for i in range(B)], bert, clip, bert_tok, pix,
                                     synth.make_token_mask("cuda"), logging.getLogger("test"), prompt=synth.SYNTH_PROMPT,
                                     batch_size=B, max_len=n, top_k=K, temperature=0.1, max_iter=sweeps, alpha=0.02,
                                     beta=2.0, generate_order=order)
    stats = runtime.any_engine().cert_stats()
    runtime.clear()
    return out, stats


@pytest.mark.parametrize("order", ["sequential", "shuffle"])
def test_config2_free_running_certified_equals_bf16x3(order, monkeypatch):
    """BASELINE config 2 / 3 at full size (64 images, sentence_len 10, candidate_k 200, 5 sweeps), free running through
    the public generate_caption: the default (certified) precision returns the same captions in every sweep, the
    same best list and bit-identical CLIP scores as the all-bf16x3 run -- 3 200 argmax decisions over 200
    candidates each."""
    (t3, s3), _ = _run_generate("bf16x3", 64, 10, 200, 5, order, monkeypatch)
    (tc, sc), st = _run_generate("certified", 64, 10, 200, 5, order, monkeypatch)
    assert tc == t3
    assert sc == s3
    assert st["calls"] == 50 and st["images"] == 3200 and st["rescored_candidates"] >= 3200
    # the point of the bound: almost every candidate is ruled out by the bf16 scores alone, and (nearly) no image
    # needs every candidate re-encoded
    assert st["rescored_candidates"] < 0.05 * 3200 * 200, st
    assert st["images_full"] < 0.05 * 3200, st


def test_config2_certified_lockstep_with_cpu_oracle():
    """B = 8 images, sentence_len 10, candidate_k 200, two sweeps (20 Gibbs steps, 160 decisions over 200 candidates):
    the certified engine in lockstep with the CPU oracle (the restatement pinned against the unmodified reference).
    Token ids must be identical after every step; a divergence is tolerated once, and only where the oracle's own
    top-2 fused scores are closer than 1e-4 (an fp32 tie no arithmetic can call)."""
    from oracle import conzic_oracle as orc
    B, n, K, sweeps = 8, 10, 200, 2
    eng = gc.engine("certified")
    o = orc.Oracle(gc.weights("bert"), gc.weights("clip"), synth.SynthBertTokenizer(), synth.SynthCLIPTokenizer(),
                   full_logits=False)
    o.trace = []
    tok = synth.SynthBertTokenizer()
    inp_ref = torch.tensor([tok.encode(synth.SYNTH_PROMPT + "[MASK]" * n)] * B)
    inp_dev = inp_ref.clone().cuda()
    img = torch.nn.functional.normalize(torch.randn(B, 512, generator=torch.Generator().manual_seed(33)), dim=-1)
    img_dev = img.cuda()
    tm_ref, tm_dev = synth.make_token_mask(), synth.make_token_mask("cuda")
    holds = [False] + [True] * 3 + [False] * (n + 1)
    near_ties = 0
    for it in range(sweeps):
        for ii in range(n):
            pos = 4 + ii
            with torch.no_grad():
                cur, _ = o.step(inp_ref, img, tm_ref, pos, ii, n, K, 0.1, 0.02, 2.0)
            cr, _, _ = eng.gibbs_step(inp_dev, tm_dev, img_dev, pos, ii == n - 1, K, 0.1, 0.02, 2.0, sum(holds[:pos]),
                                      sum(holds[pos + 1:]))
            holds[pos] = True
            got = inp_dev.cpu()
            if not torch.equal(got, inp_ref):
                t = o.trace[-1]
                top2 = t["final"].topk(2, dim=1).values
                bad = (got[:, pos] != inp_ref[:, pos])
                assert float((top2[:, 0] - top2[:, 1])[bad].max()) < 1e-4, "winner differs from the oracle's off a tie"
                near_ties += int(bad.sum())
                inp_dev.copy_(inp_ref)
            else:
                np.testing.assert_allclose(cr.cpu().numpy(), np.array(cur, dtype=np.float32), rtol=0, atol=2e-5)
            o.trace.clear()
    assert near_ties <= 1


def test_run_py_cli_synthetic_writes_reference_result_layout(tmp_path, monkeypatch):
    """run.py end to end on synthetic inputs: per-iteration JSON files + best_clipscore.json keyed by image id
    (run.py:194-222), captions identical to a direct generate_caption call with the same seed."""
    import json
    import os
    from conzic_b200 import cli, runtime
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("CONZIC_PRECISION", "bf16x3")
    runtime.clear()
    argv = ["--synthetic", "--synthetic_images", "5", "--batch_size", "2", "--run_type", "caption", "--order", "shuffle",
            "--sentence_len", "4", "--candidate_k", "8", "--num_iterations", "2", "--samples_num", "1",
            "--prompt", synth.SYNTH_PROMPT, "--results_dir", str(tmp_path / "results")]
    dirs = cli.run_main(argv)
    assert len(dirs) == 1
    files = sorted(os.listdir(dirs[0]))
    assert files == ["best_clipscore.json", "iter_0.json", "iter_1.json"]
    it1 = json.load(open(os.path.join(dirs[0], "iter_1.json")))
    assert sorted(it1) == ["synthetic0", "synthetic1", "synthetic2", "synthetic3"]  # drop_last: 5 images, batches of 2
    assert all(isinstance(v, str) and v.startswith("w3746 w1997 w1037") for v in it1.values())
    assert "caption_shuffle_len4_topk8_alpha0.020_beta2.000_gamma5.000_lmTemp0.100" in dirs[0]
    runtime.clear()


def test_demo_py_cli_synthetic(tmp_path, monkeypatch):
    from conzic_b200 import cli, runtime
    monkeypatch.chdir(tmp_path)
    runtime.clear()
    texts, scores = cli.demo_main(["--synthetic", "--run_type", "controllable", "--sentiment_type", "negative",
                                   "--order", "sequential", "--sentence_len", "4", "--candidate_k", "8",
                                   "--num_iterations", "2", "--samples_num", "1", "--prompt", synth.SYNTH_PROMPT])
    assert len(texts) == 3 and len(scores) == 3 and len(texts[0]) == 1
    runtime.clear()
