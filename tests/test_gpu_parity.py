"""GPU parity suite (-m gpu): every CUDA kernel and the whole path, called through the C ABI
(include/b200pf.h via asr-2pass_b200/capi.py), against the CPU oracle on the same seeded inputs.

Tolerances (stated once, used below):
  * integer / index work (CIF fire frames given alphas, token embeddings given alphas, argmax given logits,
    frame counts, batching)                                              : bit exact
  * fbank, LFR+CMVN                                                       : <= 1e-4 absolute (north_star)
  * GEMM / attention / LayerNorm / FSMN kernels vs a reference fed the SAME 16-bit-rounded operands : <= 4e-3
    relative (one 16-bit output rounding), fp32-output GEMMs <= 1e-5.  Every kernel test runs in both operand
    formats of the engine (fixture `prec`: IEEE fp16 = the default, bf16 = north_star's literal format).
  * encoder output and logits vs the plain fp32 oracle                    : <= 1e-2 relative (north_star) in the
    default fp16 operand format.  bf16 operands reach 1e-2 on the encoder output but only 2-4e-2 on the logits
    (operand rounding alone, in fp32 arithmetic on the CPU, moves them by 1.4-1.8e-2: oracle emulate_bf16);
    tests/test_gpu_parity_fullsize.py keeps that mode covered.
  * token ids / fire frames vs the fp32 oracle: equal, except where the oracle itself is at a tie:
    a fire may move by one frame only if the oracle's integrate value is within the accumulated alpha deviation of
    the threshold; a token id may differ only if the oracle's top-2 logit gap at that row is within twice the
    logit tolerance.  Mismatch RATES are bounded in tests/test_gpu_parity_fullsize.py.
"""
import os

import numpy as np
import pytest

from oracle import frontend as F
from oracle import paraformer_ref as R

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


_PREC = ["fp16"]   # 16-bit operand format the op_* entry points currently run in (fixture `prec`)


def bf(x):
    """Round to the 16-bit operand format under test (the name is historical: fp16 is the default format, bf16 the other)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(x, np.float32))
    return (t.half() if _PREC[0] == "fp16" else t.bfloat16()).float().numpy()


@pytest.fixture(params=["fp16", "bf16"])
def prec(request, capi):
    """Every kernel test runs in both operand formats of the engine (include/b200pf.h B200PF_PREC_*)."""
    capi.op_set_precision(request.param)
    _PREC[0] = request.param
    yield request.param
    capi.op_set_precision("fp16")
    _PREC[0] = "fp16"


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


# ---------------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 256, 64), (128, 256, 512), (300, 512, 560), (1000, 1536, 512), (2500, 2048, 512),
                                   (513, 512, 2048), (129, 1024, 512)])
def test_gemm_tcgen05(capi, gpu, shape, prec):
    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    add = rng.standard_normal((M, N)).astype(np.float32)
    res = rng.standard_normal((M, N)).astype(np.float32)
    ref = bf(A) @ bf(W).T
    assert rel(capi.op_gemm(A, W), ref) <= 1e-5
    assert rel(capi.op_gemm(A, W, general=True), ref) <= 1e-5
    full = np.maximum(ref + bias, 0) + bf(add) + res
    assert rel(capi.op_gemm(A, W, bias=bias, add=add, res=res, relu=1), full) <= 1e-5                 # in place: TMA reduce-add
    assert rel(capi.op_gemm(A, W, bias=bias, add=add, res=res, relu=1, general=True), full) <= 1e-5   # separate residual buffer
    assert np.array_equal(capi.op_gemm(A, W, bias=bias, add=add, res=res, relu=1), capi.op_gemm(A, W, bias=bias, add=add, res=res, relu=1, general=True))
    assert rel(capi.op_gemm(A, W, bias=bias, res=res, relu=2, out_bf16=True), bf(np.maximum(ref + bias + res, 0))) <= 4e-3


def test_gemm_vocab_argmax_first_max_wins(capi, gpu, prec):
    rng = np.random.default_rng(7)
    M, N, K = 257, 8404, 512
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    W[4000] = W[17]           # exact ties between two vocabulary rows: the lower index must win (util.cpp:63-74)
    W[8403] = W[17]
    bias = np.zeros(N, np.float32)
    out, am = capi.op_gemm(A, W, bias=bias, argmax=True)
    assert rel(out, bf(A) @ bf(W).T) <= 1e-5
    expect = np.array([F.find_max(out[i])[1] for i in range(M)])
    assert np.array_equal(am, expect)
    assert not np.any(am == 4000) and not np.any(am == 8403)


def test_conv3_as_shifted_gemm(capi, gpu, prec):
    import torch
    rng = np.random.default_rng(1)
    lens = [50, 130, 7, 1]
    M = sum(lens) + len(lens)
    X = np.zeros((M, 512), np.float32)
    r, segs = 0, []
    for L in lens:
        X[r:r + L] = rng.standard_normal((L, 512))
        segs.append((r, L))
        r += L + 1
    w = (rng.standard_normal((512, 512, 3)) / np.sqrt(1536)).astype(np.float32)
    b = rng.standard_normal(512).astype(np.float32)
    out = capi.op_conv3(X, np.ascontiguousarray(w.transpose(0, 2, 1).reshape(512, 1536)), b)
    for (r0, L) in segs:
        xs = torch.from_numpy(bf(X[r0:r0 + L])).t()[None]
        ref = torch.nn.functional.conv1d(torch.nn.functional.pad(xs, (1, 1)), torch.from_numpy(bf(w)), torch.from_numpy(b))[0].t().numpy()
        assert rel(out[r0:r0 + L], ref) <= 1e-5


@pytest.mark.parametrize("D", [512, 560, 2048])
def test_layernorm(capi, gpu, D, prec):
    import torch
    rng = np.random.default_rng(D)
    x = (rng.standard_normal((77, D)) * 3 + 1).astype(np.float32)
    g = rng.standard_normal(D).astype(np.float32)
    b = rng.standard_normal(D).astype(np.float32)
    ln = lambda t: torch.nn.functional.layer_norm(torch.from_numpy(t), (D,), torch.from_numpy(g), torch.from_numpy(b), 1e-12).numpy()
    o32, o16 = capi.op_layernorm(x, g, b)
    assert np.abs(o32 - ln(x)).max() <= 1e-5
    assert rel(o16, ln(x)) <= 4e-3
    o32, _ = capi.op_layernorm(x, g, b, in_bf16=True)
    assert np.abs(o32 - ln(bf(x))).max() <= 1e-5


def _attn_ref(q, k, v, q_off, q_len, kv_off, kv_len, H=4):
    out = np.zeros_like(q)
    qb, kb, vb = bf(q), bf(k), bf(v)
    for s in range(len(q_len)):
        for h in range(H):
            c = slice(h * 128, (h + 1) * 128)
            qs, ks, vs = qb[q_off[s]:q_off[s] + q_len[s], c], kb[kv_off[s]:kv_off[s] + kv_len[s], c], vb[kv_off[s]:kv_off[s] + kv_len[s], c]
            sc = (qs @ ks.T) * (128 ** -0.5)
            p = np.exp(sc - sc.max(1, keepdims=True))
            out[q_off[s]:q_off[s] + q_len[s], c] = (p / p.sum(1, keepdims=True)) @ vs
    return out


@pytest.mark.parametrize("case", [([33], [33]), ([1], [1]), ([64], [64]), ([65], [65]), ([128], [128]), ([129], [129]), ([167], [167]),
                                  ([200, 1, 64, 129], [200, 1, 64, 129]), ([40, 90], [83, 167]), ([1000], [1000])])
@pytest.mark.parametrize("impl", [0, 1])
def test_attention(capi, gpu, case, impl, prec):
    q_lens, kv_lens = case
    rng = np.random.default_rng(sum(q_lens) + 13 * sum(kv_lens))
    q_off = np.concatenate([[0], np.cumsum(q_lens)[:-1]]).astype(np.int32)
    kv_off = np.concatenate([[0], np.cumsum(np.asarray(kv_lens) + 1)[:-1]]).astype(np.int32)  # one gap row per segment
    q = rng.standard_normal((int(sum(q_lens)), 512)).astype(np.float32)
    k = rng.standard_normal((int(sum(kv_lens) + len(kv_lens)), 512)).astype(np.float32)
    v = rng.standard_normal((int(sum(kv_lens) + len(kv_lens)), 512)).astype(np.float32)
    ref = _attn_ref(q, k, v, q_off, q_lens, kv_off, kv_lens)
    out = capi.op_attention(q, k, v, q_off, q_lens, kv_off, kv_lens, impl=impl)
    assert rel(out, ref) <= 8e-3  # bf16 P (product kernel) + bf16 output rounding


@pytest.mark.parametrize("impl", [0, 1])
def test_attention_large_scores_force_the_running_maximum_to_move(capi, gpu, impl, prec):
    """Scores grow along the keys (each 64-key block's maximum is far above the previous one), so the single-pass kernel
    must take its rare path: rescale O and l in TMEM.  The tcgen05 kernel and the CUDA-core cross-check against the fp32 softmax."""
    rng = np.random.default_rng(21)
    q_lens, kv_lens = [130, 70, 200], [300, 129, 640]
    q_off = np.concatenate([[0], np.cumsum(q_lens)[:-1]]).astype(np.int32)
    kv_off = np.concatenate([[0], np.cumsum(np.asarray(kv_lens) + 1)[:-1]]).astype(np.int32)
    q = rng.standard_normal((int(sum(q_lens)), 512)).astype(np.float32)
    k = rng.standard_normal((int(sum(kv_lens) + len(kv_lens)), 512)).astype(np.float32)
    v = rng.standard_normal(k.shape).astype(np.float32)
    q[:, :4] = 6.0                                                    # a shared direction ...
    for s in range(len(kv_lens)):
        ramp = np.linspace(-8.0, 8.0, kv_lens[s]).astype(np.float32)  # ... along which the keys ramp up: +~19 in log2 per 64 keys
        k[kv_off[s]:kv_off[s] + kv_lens[s], :4] = ramp[:, None]
    q[5] *= 0.01                                                      # one flat row inside a tile whose neighbours move
    ref = _attn_ref(q, k, v, q_off, q_lens, kv_off, kv_lens)
    out = capi.op_attention(q, k, v, q_off, q_lens, kv_off, kv_lens, impl=impl)
    assert np.isfinite(out).all()
    assert rel(out, ref) <= 8e-3


def test_fsmn(capi, gpu, prec):
    import torch
    rng = np.random.default_rng(4)
    lens = [1, 5, 11, 40, 300]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    x = rng.standard_normal((off[-1], 512)).astype(np.float32)
    w = (rng.standard_normal((512, 1, 11)) * 0.3).astype(np.float32)
    out = capi.op_fsmn(x, w, off)
    for s in range(len(lens)):
        xs = torch.from_numpy(bf(x[off[s]:off[s + 1]]))
        y = torch.nn.functional.conv1d(torch.nn.functional.pad(xs.t()[None], (5, 5)), torch.from_numpy(w), groups=512)[0].t() + xs
        assert rel(out[off[s]:off[s + 1]], y.numpy()) <= 4e-3


def test_cif_is_bit_exact(capi, gpu):
    import torch
    rng = np.random.default_rng(5)
    lens = [1, 3, 40, 167, 1000]
    off = np.concatenate([[0], np.cumsum(np.asarray(lens) + 1)]).astype(np.int32)
    alphas = rng.uniform(0.0, 1.0, off[-1]).astype(np.float32)
    alphas[10:20] = 0.5            # exact threshold hits: 0.5 + 0.5 >= 1.0 must fire
    hidden = rng.standard_normal((off[-1], 512)).astype(np.float32)
    for s in range(len(lens)):
        alphas[off[s + 1] - 1] = 0.45
        hidden[off[s + 1] - 1] = 0
    n_tok, fires, emb, ff = capi.op_cif(alphas, hidden, off)
    t0 = 0
    for s in range(len(lens)):
        e_ref, f_ref = R.cif(torch.from_numpy(hidden[off[s]:off[s + 1]]), torch.from_numpy(alphas[off[s]:off[s + 1]]), 1.0)
        L = e_ref.shape[0]
        assert n_tok[s] == L
        assert np.array_equal(fires[off[s]:off[s + 1]], f_ref.numpy())
        assert np.array_equal(emb[t0:t0 + L], e_ref.numpy())
        assert np.array_equal(ff[t0:t0 + L], np.where(f_ref.numpy() >= 1.0)[0])
        t0 += L


# ---------------------------------------------------------------------------------------------------
# whole path
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def small(capi, synth, gpu, tmp_path_factory):
    import torch
    d = str(tmp_path_factory.mktemp("small"))
    cfg, W, means, vars_, toks = synth.write_synthetic_model_dir(d, dict(n_enc=2, n_dec=2), seed=0, jitter_ln=True)
    pc = R.PfConfig.from_dict(cfg)
    eng = capi.Engine(d, max_rows=2048, max_segments=64)
    eng.set_option("taps", 1)
    return dict(eng=eng, W={k: torch.from_numpy(v) for k, v in W.items()}, pc=pc, means=means, vars=vars_, toks=toks)


@pytest.mark.parametrize("n", [400, 559, 560, 1359, 16000, 52800])
def test_frontend_against_reference_knf_golden(capi, small, n):
    g = np.load(os.path.join(GOLD, "frontend_golden.npz"))
    fb, feats = small["eng"].frontend(g["pcm_%d" % n])
    assert fb.shape == g["fbank_%d" % n].shape
    assert np.abs(fb - g["fbank_%d" % n]).max() <= 1e-4           # vs the reference's own compiled knf
    ref_feats = F.lfr_cmvn(g["fbank_%d" % n], small["means"], small["vars"])
    assert np.abs(feats - ref_feats).max() <= 1e-4


def test_frontend_long_segment(capi, synth, small):
    pcm = synth.make_audio(960000, 5)  # 60 s: T = 1000
    fb, feats = small["eng"].frontend(pcm)
    x = pcm.astype(np.float32) / np.float32(32768)
    fbo = F.fbank(x)
    assert fb.shape == (5998, 80) and feats.shape == (1000, 560)
    assert np.abs(fb - fbo).max() <= 1e-4
    assert np.abs(feats - F.lfr_cmvn(fbo, small["means"], small["vars"])).max() <= 1e-4


def _oracle(model, pcm16):
    x = pcm16.astype(np.float32) / np.float32(32768)
    feats = F.lfr_cmvn(F.fbank(x), model["means"], model["vars"])
    return feats, R.forward(feats, model["W"], model["pc"])


def _check_segment(b, res, i, model, o, enc_tol=1e-2, logit_tol=1e-2):
    T = o["enc"].shape[0]
    assert res["lfr_frames"][i] == T
    enc = b.tap("enc", i)
    assert rel(enc, o["enc"].numpy()) <= enc_tol
    al = b.tap("alphas", i)
    assert al.shape == (T + 1,) and al[-1] == np.float32(0.45)
    assert np.abs(al - o["alphas"].numpy()).max() <= 5e-3
    s, e = res["token_offsets"][i], res["token_offsets"][i + 1]
    ids, fr = res["token_ids"][s:e], res["fire_frames"][s:e]
    assert res["token_counts"][i] == e - s
    # fires: the GPU scan applied to the GPU's alphas must equal the oracle's CIF recurrence applied to the SAME
    # alphas bit for bit (the scan is integer/index work once the alphas are fixed)
    import torch
    _, fires_self = R.cif(torch.zeros(T + 1, 1), torch.from_numpy(al), 1.0)
    assert np.array_equal(fr, np.where(fires_self.numpy() >= 1.0)[0])
    assert np.array_equal(b.tap("fires", i), fires_self.numpy())
    # against the oracle's own alphas a fire may move by one frame, and only where the oracle's integrate value is
    # closer to the threshold than the accumulated alpha deviation (alphas agree to <= 5e-3 per frame)
    fires_o = o["fires"].numpy()
    fr_o = np.where(fires_o >= 1.0)[0]
    assert abs(len(fr) - len(fr_o)) <= 1
    drift = np.abs(np.cumsum(al.astype(np.float64)) - np.cumsum(o["alphas"].numpy().astype(np.float64)))
    if len(fr) == len(fr_o):
        for a, c in zip(fr, fr_o):
            if a != c:
                lo, hi = min(a, c), max(a, c)
                assert hi - lo == 1 and min(abs(fires_o[a] - 1.0), abs(fires_o[c] - 1.0)) <= drift[:hi + 1].max() + 1e-3
        moved = int((fr != fr_o).sum())
        if len(fr) == 0:
            return
        lg_o = o["logits"].numpy()
        lg = b.tap("logits", i)
        if moved == 0:
            assert rel(lg, lg_o) <= logit_tol
        top2 = np.sort(lg_o, axis=1)[:, -2:]
        gap = top2[:, 1] - top2[:, 0]
        for j, (a, c) in enumerate(zip(ids, o["ids"])):
            if a != c and moved == 0:
                assert gap[j] <= 2.0 * logit_tol * np.abs(lg_o).max(), (i, j, gap[j])
        # the fused argmax is exact on the GPU's own logits (first maximum wins)
        assert np.array_equal(ids, [F.find_max(lg[j])[1] for j in range(len(ids))])


def test_forward_small_model_ragged_batch(capi, synth, small):
    lens = [16000, 52800, 160000, 320, 84000, 400, 0]       # includes too-short and empty segments
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pcm = np.concatenate([synth.make_audio(n, 100 + i) if n else np.zeros(0, np.int16) for i, n in enumerate(lens)])
    b = capi.Batch(small["eng"], int(offs[-1]) + 16)
    res = b.forward_s16(pcm, offs)
    assert res["token_counts"][3] == 0 and res["lfr_frames"][3] == 0     # n_fb == 0 -> "" (paraformer.cpp:477-480)
    assert res["token_counts"][6] == 0
    for i, n in enumerate(lens):
        if n >= 400:
            _, o = _oracle(small, pcm[offs[i]:offs[i + 1]])
            _check_segment(b, res, i, small, o)
    assert b.launches > 0 and b.flops > 0


def test_small_model_against_committed_golden(capi, small):
    g = np.load(os.path.join(GOLD, "frontend_golden.npz"))
    mg = np.load(os.path.join(GOLD, "model_small_golden.npz"))
    for n in (16000, 52800):
        pcm = g["pcm_%d" % n]
        b = capi.Batch(small["eng"], len(pcm) + 16)
        res = b.forward_s16(pcm, np.array([0, len(pcm)], np.int64))
        assert rel(b.tap("enc", 0), mg["enc_%d" % n].astype(np.float32)) <= 1.2e-2   # fixture stored as fp16
        assert abs(int(res["token_counts"][0]) - int(mg["token_num_%d" % n][0])) <= 1
        if res["token_counts"][0] == mg["token_num_%d" % n][0]:
            ids = res["token_ids"]
            bad = [(j, mg["top_gap_%d" % n][j]) for j in range(len(ids)) if ids[j] != mg["ids_%d" % n][j]]
            assert all(gap < 0.06 for _, gap in bad), bad


def test_peaky_cases_are_bit_exact(capi, synth, gpu, tmp_path_factory):
    """Committed cases (tests/golden/peaky_cases.json, made by tests/golden/make_peaky_cases.py from the fp32 oracle) on the
    "peaky" small model, chosen so that every row's top-1 margin (>= 0.05), every CIF integrate value (>= 0.01 from the
    threshold) and the token count clear their thresholds by far more than the CUDA path's rounding error: ids, fire frames
    and counts must be EXACTLY the oracle's, in both operand formats, alone and batched together."""
    import json
    g = json.load(open(os.path.join(GOLD, "peaky_cases.json")))
    m = g["model"]
    d = str(tmp_path_factory.mktemp("peaky"))
    cfg, W = synth.make_weights(m["cfg"], m["seed"], m["jitter_ln"])
    getattr(synth, m["transform"])(W)
    modelfile = __import__("importlib").import_module("asr-2pass_b200.modelfile")
    modelfile.write_model_dir(d, cfg, W, *synth.make_cmvn(int(cfg["feat_dim"])), synth.make_tokens(int(cfg["vocab"])))
    assert len(g["cases"]) >= 5
    segs = [synth.make_audio(c["n_samples"], c["audio_seed"]) for c in g["cases"]]
    offs = np.concatenate([[0], np.cumsum([len(s) for s in segs])]).astype(np.int64)
    for prec in ("fp16", "bf16"):
        eng = capi.Engine(d, max_rows=2048, max_segments=64, prec=prec)
        b = capi.Batch(eng, int(offs[-1]) + 64)
        res = b.forward_s16(np.concatenate(segs), offs)
        for i, c in enumerate(g["cases"]):
            s, e = res["token_offsets"][i], res["token_offsets"][i + 1]
            assert res["lfr_frames"][i] == c["T"] and res["token_counts"][i] == c["L"]
            assert list(res["token_ids"][s:e]) == c["ids"], (prec, i)
            assert list(res["fire_frames"][s:e]) == c["fire_frames"], (prec, i)
            r1 = b.forward_s16(segs[i], np.array([0, len(segs[i])], np.int64))      # alone: the same
            assert list(r1["token_ids"]) == c["ids"] and list(r1["fire_frames"]) == c["fire_frames"]
        b.close()
        eng.close()


def test_batch_invariance_and_input_formats(capi, synth, small):
    """A segment's result does not depend on what it is batched with (the reference's ORT path is batch-1),
    nor on whether PCM arrives as int16 or as the reference's float/32768."""
    lens = [52800, 16000, 160000]
    segs = [synth.make_audio(n, 300 + i) for i, n in enumerate(lens)]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    b = capi.Batch(small["eng"], int(offs[-1]) + 16)
    res = b.forward_s16(np.concatenate(segs), offs)
    enc_b = [b.tap("enc", i) for i in range(3)]
    for i, s in enumerate(segs):
        r1 = b.forward_s16(s, np.array([0, len(s)], np.int64))
        assert np.array_equal(r1["token_ids"], res["token_ids"][res["token_offsets"][i]:res["token_offsets"][i + 1]])
        assert np.array_equal(r1["fire_frames"], res["fire_frames"][res["token_offsets"][i]:res["token_offsets"][i + 1]])
        assert np.array_equal(b.tap("enc", 0), enc_b[i])           # bit identical
    rf = b.forward_f32([s.astype(np.float32) / np.float32(32768) for s in segs])
    assert np.array_equal(rf["token_ids"], res["token_ids"]) and np.array_equal(rf["fire_frames"], res["fire_frames"])
    r2 = b.forward_s16(np.concatenate(segs), offs)                 # run-to-run determinism
    assert np.array_equal(r2["token_ids"], res["token_ids"])
    # float input that is NOT int16 / 32768 (e.g. resampled audio) cannot take the exact int16 staging path: it is copied as
    # float, and the front end follows the reference's float arithmetic (paraformer.cpp:312-314) to the same 1e-4
    eng = small["eng"]
    xf = [(s.astype(np.float32) / np.float32(32768)) * np.float32(0.37) + np.float32(1e-5) for s in segs[:2]]
    eng.set_option("taps", 1)
    rf2 = b.forward_f32(xf)
    for i, x in enumerate(xf):
        fb = F.fbank(x)
        assert np.abs(b.tap("fbank", i) - fb).max() <= 1e-4
    assert rf2["token_counts"][0] > 0


def test_capacity_and_argument_errors(capi, synth, small):
    eng = small["eng"]
    b = capi.Batch(eng, 16000 * 200)
    long_pcm = np.zeros(16000 * 130, np.int16)                     # 130 s -> 2167 rows > max_rows 2048
    with pytest.raises(capi.B200PFError, match="max_rows"):
        b.forward_s16(long_pcm, np.array([0, len(long_pcm)], np.int64))
    with pytest.raises(capi.B200PFError, match="max_samples"):
        capi.Batch(eng, 1000).forward_s16(np.zeros(2000, np.int16), np.array([0, 2000], np.int64))
    with pytest.raises(capi.B200PFError, match="monotone"):
        b.forward_s16(np.zeros(2000, np.int16), np.array([0, 1500, 1000], np.int64))
    r = b.forward_s16(np.zeros(0, np.int16), np.array([0], np.int64))   # empty batch
    assert r["n_tokens"] == 0


def test_batch_outliving_its_engine_is_refused_not_a_crash(capi, synth, gpu, tmp_path_factory):
    """b200pf_engine_destroy orphans the batches that are still alive (include/b200pf.h): their entry points return an error and
    destroying them later only frees the handle.  Driven through the raw C ABI -- the Python wrapper closes batches first."""
    import ctypes as C
    d = str(tmp_path_factory.mktemp("orphan"))
    synth.write_synthetic_model_dir(d, dict(n_enc=1, n_dec=1))
    eng = capi.Engine(d, max_rows=2048, max_segments=16)
    b = capi.Batch(eng, 48000)
    pcm = synth.make_audio(32000, 3)
    offs = np.array([0, 32000], np.int64)
    assert b.forward_s16(pcm, offs)["n_tokens"] >= 0
    L = capi.lib()
    L.b200pf_engine_destroy(eng.h)          # engine first: the batch is now an orphan
    eng.h = C.c_void_p()
    eng._batches = []
    with pytest.raises(capi.B200PFError):
        b.forward_s16(pcm, offs)
    b.close()                               # frees the handle only; must not touch the engine's memory


def test_forward_full_paraformer_large(capi, synth, gpu, tmp_path_factory):
    """The named architecture (50 + 16 layers, 215.8 M parameters): config[0]'s 10 s segment plus a short and a
    long one, against the fp32 oracle."""
    import torch
    d = str(tmp_path_factory.mktemp("full"))
    cfg, W, means, vars_, toks = synth.write_synthetic_model_dir(d, None, seed=0)
    pc = R.PfConfig()
    model = dict(W={k: torch.from_numpy(v) for k, v in W.items()}, pc=pc, means=means, vars=vars_)
    eng = capi.Engine(d, max_rows=2048, max_segments=16)
    eng.set_option("taps", 1)
    lens = [160000, 32000, 320000]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pcm = np.concatenate([synth.make_audio(n, 1234 + i) for i, n in enumerate(lens)])
    b = capi.Batch(eng, int(offs[-1]) + 16)
    res = b.forward_s16(pcm, offs)
    assert list(res["lfr_frames"]) == [167, 33, 333]
    for i in range(3):
        _, o = _oracle(model, pcm[offs[i]:offs[i + 1]])
        assert abs(int(res["token_counts"][i]) - o["token_num"]) <= 1
        _check_segment(b, res, i, model, o)        # 1e-2 on encoder output and logits (north_star)
    # 574 launches: 2 front end + 50 x 8 encoder + 6 predictor + 16 x 10 decoder + 6 tail (the feed-forward LayerNorm of the 17
    # decoder FFNs is folded into the GEMMs around it, engine option ffn_ln_fold)
    assert b.launches == 2 + 50 * 8 + 6 + 16 * 10 + 6
    # the same forward with the LayerNorm as its own pass: same bar against the oracle, (nearly) the same ids
    eng.set_option("ffn_ln_fold", 0)
    res0 = b.forward_s16(pcm, offs)
    assert b.launches == 2 + 50 * 8 + 6 + 16 * 11 + 7
    for i in range(3):
        _, o = _oracle(model, pcm[offs[i]:offs[i + 1]])
        _check_segment(b, res0, i, model, o)
    assert list(res0["token_counts"]) == list(res["token_counts"])
    assert np.mean(np.asarray(res0["token_ids"]) != np.asarray(res["token_ids"])) <= 0.02
    eng.close()


def test_recording_longer_than_the_position_table_as_one_segment(capi, synth, gpu, tmp_path_factory):
    """FunASRInferBuffer decodes a whole recording as ONE segment (no VAD on that API, funasrruntime.cpp:57-114).  Beyond the 2048
    rows of the precomputed table the position encoding is evaluated in the kernel: a 135 s segment (T = 2250) agrees with the
    oracle like any other, and a recording that exceeds the engine's capacity is a real error, not an empty transcript."""
    import torch
    d = str(tmp_path_factory.mktemp("long"))
    cfg, W, means, vars_, toks = synth.write_synthetic_model_dir(d, dict(n_enc=2, n_dec=2), seed=0, jitter_ln=True)
    model = dict(W={k: torch.from_numpy(v) for k, v in W.items()}, pc=R.PfConfig.from_dict(cfg), means=means, vars=vars_)
    eng = capi.Engine(d, max_rows=4096, max_segments=8)
    eng.set_option("taps", 1)
    pcm = synth.make_audio(16000 * 135, 9)
    b = capi.Batch(eng, len(pcm) + 16)
    res = b.forward_s16(pcm, np.array([0, len(pcm)], np.int64))
    assert res["lfr_frames"][0] == 2250
    _, o = _oracle(model, pcm)
    _check_segment(b, res, 0, model, o)
    b.close()
    eng.close()
    text = capi.funasr_infer(d, pcm16=pcm, max_rows=4096)                  # fits: decoded as one segment
    assert len(text) > 100
    with pytest.raises(capi.B200PFError):                                  # does not fit: an error, not ""
        capi.funasr_infer(d, pcm16=pcm, max_rows=1024)


def test_max_length_segment_full_size_properties(capi, synth, small):
    """60 s (vad_max_len) segment, T = 1000: size-independent properties instead of an oracle comparison."""
    pcm = synth.make_audio(960000, 77)
    b = capi.Batch(small["eng"], len(pcm) + 16)
    res = b.forward_s16(pcm, np.array([0, len(pcm)], np.int64))
    assert res["lfr_frames"][0] == 1000
    al, fires = b.tap("alphas", 0), b.tap("fires", 0)
    L = int(res["token_counts"][0])
    assert L == int((fires >= 1.0).sum())                              # tokens == fires
    assert abs(L - np.floor(al.astype(np.float64).sum())) <= 1          # token_num = floor(sum alpha)
    fr = res["fire_frames"]
    assert np.all(np.diff(fr) > 0) and fr[-1] <= 1000                    # fire frames strictly increasing
    assert np.all((res["token_ids"] >= 0) & (res["token_ids"] < 8404))
    emb = b.tap("embeds", 0)
    assert emb.shape == (L, 512) and np.isfinite(emb).all()


def test_microbatcher_results_equal_direct_calls(capi, synth, small, tmp_path_factory):
    """batch-1 Forward calls of many connections merged into batched forwards give, per
    segment, exactly the string a direct call gives (the engine is batch invariant)."""
    import threading
    d = str(tmp_path_factory.mktemp("mb"))
    synth.write_synthetic_model_dir(d, dict(n_enc=2, n_dec=2), seed=0, jitter_ln=True)
    h = capi.OfflineHandle(d, max_rows=4096, max_segments=64, batch_size=64)
    segs = [synth.make_audio(int(n), 4000 + i).astype(np.float32) / np.float32(32768)
            for i, n in enumerate([16000, 52800, 33000, 8000, 120000, 20000, 64000, 300, 48000, 25000, 90000, 16000])]
    direct = [h.model_forward([s])[0] for s in segs]
    mb = capi.MicroBatcher(h, max_wait_us=30000, max_batch=64, max_rows=4096)
    out = [None] * len(segs)

    def call(i):
        out[i] = mb.forward(segs[i])

    th = [threading.Thread(target=call, args=(i,)) for i in range(len(segs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert out == direct
    assert out[7] == ""                                            # too short -> "" also through the batcher
    st = mb.stats()
    assert st["segments"] == len(segs) and st["batches"] <= 3 and st["max_batch_seen"] >= 4
    mb.close()
    h.close()


def test_logprob_topk_kernel(capi, gpu):
    """Pruned posteriors: logsumexp + top-k log-softmax with the (value desc, index asc) order."""
    rng = np.random.default_rng(17)
    x = (rng.standard_normal((37, 8404)) * 3).astype(np.float32)
    x[3, 100] = x[3, 7000] = x[3].max() + 1.0            # exact tie at the top: the lower index comes first
    x[5, :] = 0.25                                        # a constant row: ids 0..k-1
    lse, lp, ids = capi.op_logprob_topk(x, 16)
    lse_o, lp_o, ids_o = R.logprob_topk(x, 16)
    assert np.array_equal(ids, ids_o)
    assert np.abs(lse - lse_o).max() <= 1e-5 and np.abs(lp - lp_o).max() <= 1e-5
    assert list(ids[3, :2]) == [100, 7000] and list(ids[5]) == list(range(16))
    assert np.all(np.diff(lp, axis=1) <= 0)


def test_forward_returns_pruned_posteriors(capi, synth, small):
    eng = small["eng"]
    pcm = synth.make_audio(52800, 321)
    b = capi.Batch(eng, len(pcm) + 16)
    eng.set_option("logprob_topk", 8)
    try:
        res = b.forward_s16(pcm, np.array([0, len(pcm)], np.int64))
        lg = b.tap("logits", 0)
        lse_o, lp_o, ids_o = R.logprob_topk(lg, 8)
        assert res["topk_ids"].shape == (res["n_tokens"], 8)
        assert np.array_equal(res["topk_ids"], ids_o)
        assert np.array_equal(res["topk_ids"][:, 0], res["token_ids"])            # entry 0 is the greedy token
        assert np.abs(res["topk_logprob"] - lp_o).max() <= 1e-5 and np.abs(res["token_lse"] - lse_o).max() <= 1e-5
    finally:
        eng.set_option("logprob_topk", 0)
    res = b.forward_s16(pcm, np.array([0, len(pcm)], np.int64))
    assert "topk_ids" not in res
