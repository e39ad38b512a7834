"""GPU parity suite for config 3 (-m gpu): the timestamp predictor (CifPredictorV3 upsample head -> us_alphas /
us_cif_peak, paraformer.cpp:549-563), the contextual decoder (hw_emb input, paraformer.cpp:515-531) and the hotword
compiler (Embedding + LSTM, paraformer.cpp:592-693), through the C ABI against the CPU oracle.

Tolerances:
  * cif_wo_hidden given the GPU's own alphas                         : bit exact (sequential fp32 recurrence)
  * LSTM kernel vs the oracle recurrence fed the same bf16 operands  : <= 2e-2 absolute on h in (-1, 1) (the hidden
    state is fed back in bf16 on both sides; the orders of the 512-term dot products differ)
  * us_alphas vs the fp32 oracle                                     : <= 3e-2 relative to the largest alpha
  * a timestamp peak may move by one upsampled frame (20 ms) where the oracle's running sum is within the
    accumulated alpha deviation of the threshold
  * logits of the contextual decoder                                 : same 1e-2 as the plain decoder
"""
import os

import numpy as np
import pytest

from oracle import frontend as F
from oracle import paraformer_ref as R
from oracle import postproc_ref as P

pytestmark = pytest.mark.gpu


def bf(x):
    """Round to the engine's default 16-bit operand format (IEEE fp16; include/b200pf.h B200PF_PREC_FP16)."""
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).half().float().numpy()


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _lstm_weights(rng, n_dir):
    k = 1.0 / np.sqrt(512)
    w_ih = rng.uniform(-k, k, (n_dir * 2048, 512)).astype(np.float32)
    w_hh = rng.uniform(-k, k, (n_dir * 2048, 512)).astype(np.float32)
    b_ih = rng.uniform(-k, k, n_dir * 2048).astype(np.float32)
    b_hh = rng.uniform(-k, k, n_dir * 2048).astype(np.float32)
    return w_ih, w_hh, b_ih, b_hh


def _oracle_lstm(x, w_ih, w_hh, b_ih, b_hh, d, emu):
    import torch
    W = {"l.weight_ih_l0": torch.from_numpy(w_ih[d * 2048:(d + 1) * 2048]), "l.weight_hh_l0": torch.from_numpy(w_hh[d * 2048:(d + 1) * 2048]),
         "l.bias_ih_l0": torch.from_numpy(b_ih[d * 2048:(d + 1) * 2048]), "l.bias_hh_l0": torch.from_numpy(b_hh[d * 2048:(d + 1) * 2048])}
    return R.lstm(torch.from_numpy(x), W, "l", "", reverse=(d == 1), emu=emu).numpy()


@pytest.mark.parametrize("n_dir", [1, 2])
def test_lstm_kernel(capi, gpu, n_dir):
    """Ragged sequences in two cluster groups (> 32 sequences), gaps between them, both directions."""
    rng = np.random.default_rng(11 + n_dir)
    lens = [1, 2, 3, 10, 37, 64, 5, 150] + [int(v) for v in rng.integers(1, 40, 30)]
    offs, r = [], 0
    for L in lens:
        offs.append(r)
        r += L + 2                       # rows between sequences belong to nobody and must stay untouched (zero)
    rows = r
    x = rng.standard_normal((rows, 512)).astype(np.float32)
    w_ih, w_hh, b_ih, b_hh = _lstm_weights(rng, n_dir)
    out = capi.op_lstm(x, offs, lens, w_ih, w_hh, b_ih, b_hh)
    out16 = capi.op_lstm(x, offs, lens, w_ih, w_hh, b_ih, b_hh, bf16_out=True)
    covered = np.zeros(rows, bool)
    for o, L in zip(offs, lens):
        covered[o:o + L] = True
        for d in range(n_dir):
            ref = _oracle_lstm(x[o:o + L], w_ih, w_hh, b_ih, b_hh, d, emu="fp16")
            got = out[o:o + L, d * 512:(d + 1) * 512]
            assert np.abs(got - ref).max() <= 2e-2, (L, d, np.abs(got - ref).max())
            ref32 = _oracle_lstm(x[o:o + L], w_ih, w_hh, b_ih, b_hh, d, emu=False)
            assert np.abs(got - ref32).max() <= 4e-2
            assert np.abs(out16[o:o + L, d * 512:(d + 1) * 512] - bf(got)).max() == 0      # bf16 output = rounded fp32 output
    assert np.all(out[~covered] == 0) and np.all(out16[~covered] == 0)


def test_us_peaks_scan_is_bit_exact(capi, gpu):
    import torch
    rng = np.random.default_rng(3)
    lens = [3, 99, 501, 3000, 33]
    offs = np.concatenate([[0], np.cumsum(np.asarray(lens) + 3)[:-1]]).astype(np.int32)
    rows = int(offs[-1] + lens[-1] + 3)
    a2 = rng.uniform(0.0, 0.25, rows).astype(np.float32)
    n_tok = [1, 17, 80, 499, 0]
    th = float(np.float32(1.0 - 1e-4))
    ua, up = capi.op_us_peaks(a2, offs, lens, n_tok, th)
    for o, L, nt in zip(offs, lens, n_tok):
        seg = a2[o:o + L]
        ratio = np.float32(nt) / np.float32(seg.astype(np.float64).sum())
        assert np.allclose(ua[o:o + L], seg * ratio, rtol=2e-6, atol=0)
        assert abs(float(ua[o:o + L].astype(np.float64).sum()) - nt) <= 1e-3 * max(nt, 1)
        ref = R.cif_wo_hidden(torch.from_numpy(ua[o:o + L].copy()), 1.0 - 1e-4).numpy()
        assert np.array_equal(up[o:o + L], ref)


@pytest.fixture(scope="module")
def cfg3(capi, synth, gpu, tmp_path_factory):
    import torch
    d = str(tmp_path_factory.mktemp("cfg3"))
    cfg, W, means, vars_, toks = synth.write_synthetic_model_dir(d, dict(n_enc=2, n_dec=2, timestamp=1, contextual=1), seed=3, jitter_ln=True)
    pc = R.PfConfig.from_dict(cfg)
    eng = capi.Engine(d, max_rows=2048, max_segments=64)
    eng.set_option("taps", 1)
    return dict(dir=d, eng=eng, W={k: torch.from_numpy(v) for k, v in W.items()}, pc=pc, means=means, vars=vars_, toks=toks)


def _hotword_ids(rng, n, vocab):
    ids = np.zeros((n + 1, 10), np.int32)
    lens = np.zeros(n + 1, np.int32)
    for j in range(n):
        L = int(rng.integers(2, 7))
        ids[j, :L] = rng.integers(3, vocab - 1, L)
        lens[j] = L
    ids[n, 0] = 1                       # the blank row CompileHotwordEmbedding appends (paraformer.cpp:644-647)
    lens[n] = 1
    return ids, lens


def test_hotword_embedding_against_oracle(capi, cfg3):
    rng = np.random.default_rng(0)
    ids, lens = _hotword_ids(rng, 100, cfg3["pc"].vocab)
    got = cfg3["eng"].hotword_embed(ids, lens)
    ref = R.select_hotword_rows(R.hotword_embed(ids, cfg3["W"], emu="fp16"), lens).numpy()
    assert got.shape == (101, 512)
    assert np.abs(got - ref).max() <= 2e-2
    ref32 = R.select_hotword_rows(R.hotword_embed(ids, cfg3["W"]), lens).numpy()
    assert np.abs(got - ref32).max() <= 4e-2


def test_config3_model_config_is_reported(capi, cfg3):
    c = cfg3["eng"].cfg
    assert c.timestamp == 1 and c.contextual == 1
    cfgp, _, _ = capi.model_dir_probe(cfg3["dir"])
    assert cfgp.timestamp == 1 and cfgp.contextual == 1


def test_contextual_model_without_hotwords_fails_like_the_reference(capi, synth, cfg3):
    pcm = synth.make_audio(16000, 1)
    b = capi.Batch(cfg3["eng"], 20000)
    with pytest.raises(capi.B200PFError, match="hw_emb is null"):
        b.forward_s16(pcm, np.array([0, len(pcm)], np.int64))
    with pytest.raises(capi.B200PFError, match="dimension"):
        b.set_hotwords(np.zeros((3, 100), np.float32))


def _peaks(x):
    return np.where(x > 1.0 - 1e-4)[0]


def test_forward_config3_small_model(capi, synth, cfg3):
    """Timestamp head + contextual decoder on a ragged batch against the fp32 oracle."""
    import torch
    rng = np.random.default_rng(5)
    ids_hw, lens_hw = _hotword_ids(rng, 100, cfg3["pc"].vocab)
    hw = cfg3["eng"].hotword_embed(ids_hw, lens_hw)                     # 101 x 512, as config 3 asks
    lens = [16000, 52800, 160000, 320, 84000]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pcm = np.concatenate([synth.make_audio(n, 500 + i) for i, n in enumerate(lens)])
    b = capi.Batch(cfg3["eng"], int(offs[-1]) + 16)
    b.set_hotwords(hw)
    res = b.forward_s16(pcm, offs)
    assert res["token_counts"][3] == 0 and res["us_offsets"][4] == res["us_offsets"][3]
    for i, n in enumerate(lens):
        if n < 400:
            continue
        x = pcm[offs[i]:offs[i + 1]].astype(np.float32) / np.float32(32768)
        feats = F.lfr_cmvn(F.fbank(x), cfg3["means"], cfg3["vars"])
        o = R.forward(feats, cfg3["W"], cfg3["pc"], hw_emb=torch.from_numpy(hw))
        T = feats.shape[0]
        ua = res["us_alphas"][res["us_offsets"][i]:res["us_offsets"][i + 1]]
        up = res["us_peaks"][res["us_offsets"][i]:res["us_offsets"][i + 1]]
        assert ua.shape == (3 * T,) and up.shape == (3 * T,)
        cnt = int(res["token_counts"][i])
        # given its own alphas the scan is exact, and the alphas sum to the token count
        assert np.array_equal(up, R.cif_wo_hidden(torch.from_numpy(ua.copy()), 1.0 - 1e-4).numpy())
        assert abs(float(ua.astype(np.float64).sum()) - cnt) <= 1e-3 * max(cnt, 1)
        if cnt != o["token_num"]:
            assert abs(cnt - o["token_num"]) == 1       # a CIF fire at a tie moved (see test_gpu_parity tolerances)
            continue
        ua_o, up_o = o["us_alphas"].numpy(), o["us_peaks"].numpy()
        assert rel(ua, ua_o) <= 3e-2
        pk, pk_o = _peaks(up), _peaks(up_o)
        assert abs(len(pk) - len(pk_o)) <= 1
        if len(pk) == len(pk_o):
            drift = np.abs(np.cumsum(ua.astype(np.float64)) - np.cumsum(ua_o.astype(np.float64)))
            for a, c in zip(pk, pk_o):
                if a != c:
                    lo, hi = min(a, c), max(a, c)
                    assert hi - lo == 1 and min(abs(up_o[a] - 1.0), abs(up_o[c] - 1.0)) <= drift[:hi + 1].max() + 2e-3
        # contextual decoder
        s, e = res["token_offsets"][i], res["token_offsets"][i + 1]
        ids = res["token_ids"][s:e]
        fr = res["fire_frames"][s:e]
        fr_o = np.where(o["fires"].numpy() >= 1.0)[0]
        lg = b.tap("logits", i)
        assert np.array_equal(ids, [F.find_max(lg[j])[1] for j in range(len(ids))])
        if np.array_equal(fr, fr_o):
            lg_o = o["logits"].numpy()
            assert rel(lg, lg_o) <= 1e-2
            top2 = np.sort(lg_o, axis=1)[:, -2:]
            gap = top2[:, 1] - top2[:, 0]
            for j, (a, c) in enumerate(zip(ids, o["ids"])):
                assert a == c or gap[j] < 0.06, (i, j, gap[j])


def test_config3_against_committed_golden(capi, cfg3):
    """Committed fixtures (tests/golden/model_cfg3_golden.npz, made by make_golden.py from the fp32 oracle)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = np.load(os.path.join(root, "tests", "golden", "frontend_golden.npz"))
    c3 = np.load(os.path.join(root, "tests", "golden", "model_cfg3_golden.npz"))
    hw = cfg3["eng"].hotword_embed(c3["hw_ids"], c3["hw_len"])
    assert np.abs(hw - c3["hw_emb"]).max() <= 4e-2
    for n in (16000, 52800):
        pcm = g["pcm_%d" % n]
        b = capi.Batch(cfg3["eng"], len(pcm) + 16)
        b.set_hotwords(c3["hw_emb"])
        res = b.forward_s16(pcm, np.array([0, len(pcm)], np.int64))
        assert abs(int(res["token_counts"][0]) - int(c3["token_num_%d" % n][0])) <= 1
        if res["token_counts"][0] == c3["token_num_%d" % n][0]:
            assert rel(res["us_alphas"], c3["us_alphas_%d" % n]) <= 3e-2
            assert abs(len(_peaks(res["us_peaks"])) - len(_peaks(c3["us_peaks_%d" % n]))) <= 1
            ids = res["token_ids"]
            bad = [(j, c3["top_gap_%d" % n][j]) for j in range(len(ids)) if ids[j] != c3["ids_%d" % n][j]]
            assert all(gap < 0.06 for _, gap in bad), bad


def test_hotwords_change_the_logits_and_batch_invariance(capi, synth, cfg3):
    rng = np.random.default_rng(9)
    ids_hw, lens_hw = _hotword_ids(rng, 7, cfg3["pc"].vocab)
    hw = cfg3["eng"].hotword_embed(ids_hw, lens_hw)
    segs = [synth.make_audio(n, 700 + i) for i, n in enumerate([52800, 16000])]
    offs = np.concatenate([[0], np.cumsum([len(s) for s in segs])]).astype(np.int64)
    b = capi.Batch(cfg3["eng"], int(offs[-1]) + 16)
    b.set_hotwords(hw)
    res = b.forward_s16(np.concatenate(segs), offs)
    lg = [b.tap("logits", i) for i in range(2)]
    for i, s in enumerate(segs):                                         # alone == inside a batch, bit for bit
        r1 = b.forward_s16(s, np.array([0, len(s)], np.int64))
        assert np.array_equal(b.tap("logits", 0), lg[i])
        assert np.array_equal(r1["us_peaks"], res["us_peaks"][res["us_offsets"][i]:res["us_offsets"][i + 1]])
    b.set_hotwords(hw[::-1].copy() * 3.0)                                # other hotwords -> other bias -> other logits
    b.forward_s16(np.concatenate(segs), offs)
    assert not np.array_equal(b.tap("logits", 0), lg[0])


def test_host_forward_with_hotwords_and_timestamps(capi, synth, cfg3):
    """funasr::Model seam: CompileHotwordEmbedding -> Forward(hw_emb) -> "text | b, e,b, e" (PostProcess format),
    and FunOfflineInferBuffer's "[[b,e],...]" stitching, against the Python restatement fed the GPU's own outputs."""
    toks = cfg3["toks"]
    h = capi.OfflineHandle(cfg3["dir"], max_rows=2048, max_segments=64, batch_size=8)
    hotwords = " ".join(toks[10 + 3 * k] + toks[40 + k] + toks[90 + 2 * k] for k in range(20)) + " notinvocab " + toks[7]
    emb = h.compile_hotwords(hotwords)
    assert emb.shape == (22, 512)                                        # 20 + 1 single-char word + blank; the OOV word is dropped
    ids = np.zeros((22, 10), np.int32)
    lens = np.zeros(22, np.int32)
    for k in range(20):
        ids[k, :3] = [10 + 3 * k, 40 + k, 90 + 2 * k]
        lens[k] = 3
    ids[20, 0], lens[20] = 7, 1
    ids[21, 0], lens[21] = 1, 1
    assert np.array_equal(emb, cfg3["eng"].hotword_embed(ids, lens))
    segs = [synth.make_audio(n, 900 + i) for i, n in enumerate([52800, 16000, 200])]
    out = h.model_forward([s.astype(np.float32) / np.float32(32768) for s in segs], hw_emb=emb)
    assert out[2] == ""                                                  # too short -> "" (paraformer.cpp:477-480)
    b = capi.Batch(cfg3["eng"], 80000)
    b.set_hotwords(emb)
    vocab = P.Vocab(toks)
    for i in range(2):
        r = b.forward_s16(segs[i], np.array([0, len(segs[i])], np.int64))
        expect = P.greedy_search_text(vocab, list(r["token_ids"]), "zh-cn", list(r["us_alphas"]), list(r["us_peaks"]))
        assert out[i] == expect
        assert " | " in out[i]
    assert h.model_forward([segs[0].astype(np.float32) / np.float32(32768)], hw_emb=np.zeros((0, 512), np.float32)) == [""]   # hw_emb is null
    text, stamp = h.infer_buffer_hw(segs[0], emb)
    msg0 = out[0]
    t_ref, s_ref = P.stitch_offline([msg0], [0.0], "zh-cn")
    assert text == t_ref and stamp == s_ref and stamp.startswith("[[")
    h.close()
