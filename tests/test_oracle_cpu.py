"""CPU suite (-m "not gpu"): the oracle against the reference's golden vectors, host logic, and the C ABI's
symbol table.  No compute call goes through the CUDA library here."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import frontend as F
from oracle import postproc_ref as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "frontend_golden.npz"))


# ---------------------------------------------------------------------------------------------------
# front end: oracle restatement pinned against the reference's compiled knf (golden vectors)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [399, 400, 559, 560, 1359, 16000, 52800])
def test_fbank_oracle_matches_reference_knf_golden(gold, n):
    pcm = gold["pcm_%d" % n].astype(np.float32) / np.float32(32768)
    ref = gold["fbank_%d" % n]
    got = F.fbank(pcm)
    assert got.shape == ref.shape == (F.num_fbank_frames(n), 80)
    if ref.size:
        # both sides do the FFT in double; they differ only by float rounding of the narrowed spectrum
        assert np.abs(got - ref).max() <= 2e-5
        assert (got == ref).mean() > 0.9


def test_fbank_oracle_matches_live_reference_when_built():
    if F.ref_lib() is None:
        pytest.skip("oracle/_ref not built on this box (needs /root/reference)")
    rng = np.random.default_rng(1)
    for n in (400, 4000, 31999):
        x = (np.round(rng.standard_normal(n) * 3000).clip(-32768, 32767) / 32768).astype(np.float32)
        assert np.abs(F.fbank(x) - F.fbank_ref(x)).max() <= 2e-5


def test_rfft_known_answer(gold):
    # kaldi-native-fbank/csrc/test-rfft.cc:32-50 (values of torch.fft.rfft quoted there)
    d = np.array([1, -1, 3, 8, 20, 6, 0, 2], np.float32)
    out = F.rfft_packed(d)
    assert out[0] == 39 and out[1] == 9
    np.testing.assert_allclose(out[2], -28.1924, atol=1e-3)
    np.testing.assert_allclose(-out[3], -2.2929, atol=1e-3)
    np.testing.assert_allclose(out[4], 18, atol=1e-3)
    np.testing.assert_allclose(-out[5], 5, atol=1e-3)
    np.testing.assert_allclose(out[6], -9.8076, atol=1e-3)
    np.testing.assert_allclose(-out[7], 3.7071, atol=1e-3)
    np.testing.assert_allclose(out, gold["rfft8"], atol=1e-4)  # the reference's own Rfft on the same input


@pytest.mark.parametrize("n_fb", [1, 3, 6, 7, 12, 13, 98])
def test_lfr_cmvn_oracle_matches_literal_transcription(gold, synth, n_fb):
    means, vars_ = synth.make_cmvn()
    out = F.lfr_cmvn(gold["lfr_in_%d" % n_fb], means, vars_)
    ref = gold["lfr_out_%d" % n_fb]
    assert out.shape == ref.shape == ((n_fb + 5) // 6, 560)
    assert np.array_equal(out, ref)  # bit exact: copies + one add + one multiply


def test_frame_counts():
    for n, nfb in [(0, 0), (399, 0), (400, 1), (559, 1), (560, 2), (16000, 98), (160000, 998), (960000, 5998)]:
        assert F.num_fbank_frames(n) == nfb
    assert F.num_lfr_frames(998) == 167 and F.num_lfr_frames(5998) == 1000 and F.num_lfr_frames(1) == 1


def test_find_max_first_wins():
    v, i = F.find_max(np.array([1.0, 5.0, 5.0, -2.0], np.float32))
    assert (v, i) == (5.0, 1)
    v, i = F.find_max(np.array([-np.inf, -np.inf], np.float32))
    assert i == -1  # strict '>' never fires (util.cpp:63-74)


# ---------------------------------------------------------------------------------------------------
# graph restatement: internal consistency (parity unpinned, see oracle/paraformer_ref.py)
# ---------------------------------------------------------------------------------------------------
def test_model_oracle_reproduces_committed_taps(gold, synth):
    import torch
    from oracle import paraformer_ref as R
    mg = np.load(os.path.join(GOLD, "model_small_golden.npz"))
    cfg, W = synth.make_weights(dict(n_enc=2, n_dec=2), seed=0, jitter_ln=True)
    pc = R.PfConfig.from_dict(cfg)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    means, vars_ = synth.make_cmvn()
    for n in (16000, 52800):
        feats = F.lfr_cmvn(gold["fbank_%d" % n], means, vars_)
        o = R.forward(feats, Wt, pc)
        assert o["token_num"] == int(mg["token_num_%d" % n][0]) == o["embeds"].shape[0]
        np.testing.assert_allclose(o["enc"].numpy(), mg["enc_%d" % n].astype(np.float32), atol=5e-3)
        np.testing.assert_allclose(o["alphas"].numpy(), mg["alphas_%d" % n], atol=1e-5)
        gap = mg["top_gap_%d" % n]
        ids = np.asarray(o["ids"])
        assert np.all((ids == mg["ids_%d" % n]) | (gap < 1e-4))


def test_config3_oracle_reproduces_committed_taps(gold, synth):
    import torch
    from oracle import paraformer_ref as R
    c3 = np.load(os.path.join(GOLD, "model_cfg3_golden.npz"))
    cfg, W = synth.make_weights(dict(n_enc=2, n_dec=2, timestamp=1, contextual=1), seed=3, jitter_ln=True)
    pc = R.PfConfig.from_dict(cfg)
    assert pc.timestamp == 1 and pc.contextual == 1 and pc.smooth_factor2 == 0.25
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    hw = R.select_hotword_rows(R.hotword_embed(c3["hw_ids"], Wt), c3["hw_len"])
    np.testing.assert_allclose(hw.numpy(), c3["hw_emb"], atol=1e-5)
    means, vars_ = synth.make_cmvn()
    n = 16000
    o = R.forward(F.lfr_cmvn(gold["fbank_%d" % n], means, vars_), Wt, pc, hw_emb=hw)
    assert o["token_num"] == int(c3["token_num_%d" % n][0])
    np.testing.assert_allclose(o["us_alphas"].numpy(), c3["us_alphas_%d" % n], atol=1e-5)
    assert np.array_equal(np.where(o["us_peaks"].numpy() > 1 - 1e-4)[0], np.where(c3["us_peaks_%d" % n] > 1 - 1e-4)[0])
    ids = np.asarray(o["ids"])
    assert np.all((ids == c3["ids_%d" % n]) | (c3["top_gap_%d" % n] < 1e-4))


def test_lstm_restatement_equals_torch_nn_lstm():
    """The literal recurrence in the oracle (gate order i,f,g,o, both directions) against torch.nn.LSTM itself."""
    import torch
    from oracle import paraformer_ref as R
    torch.manual_seed(1)
    m = torch.nn.LSTM(512, 512, 1, batch_first=True, bidirectional=True)
    W = {"p." + k: v.detach() for k, v in m.state_dict().items()}
    x = torch.randn(29, 512)
    with torch.no_grad():
        ref, _ = m(x[None])
    got = torch.cat([R.lstm(x, W, "p", "", False), R.lstm(x, W, "p", "_reverse", True)], 1)
    assert float((got - ref[0]).abs().max()) <= 2e-6


def test_upsample_restatement_equals_conv_transpose1d(synth):
    """ConvTranspose1d(512,512,k=3,stride=3) written as three matmuls == torch's own op; cif_wo_hidden closed form."""
    import torch
    from oracle import paraformer_ref as R
    torch.manual_seed(2)
    ct = torch.nn.ConvTranspose1d(512, 512, 3, 3)
    enc = torch.randn(11, 512)
    with torch.no_grad():
        ref = ct(enc.t()[None])[0].t()                                   # [33, 512]
    wt, b = ct.weight.detach(), ct.bias.detach()
    up = (torch.stack([enc @ wt[:, :, j] for j in range(3)], 1) + b).reshape(33, 512)
    assert float((up - ref).abs().max()) <= 1e-5
    fires = R.cif_wo_hidden(torch.full((12,), 0.25), 1.0 - 1e-4)
    # running sum 0.25, 0.5, 0.75, 1.0 (fire, minus 0.9999) ...
    assert np.array_equal(np.where(fires.numpy() > 1 - 1e-4)[0], [3, 7, 11])
    assert abs(float(fires[4]) - (0.25 + 1e-4)) < 1e-6


def test_contextual_decoder_reduces_to_self_attention_path_when_bias_output_is_zero(synth):
    """With bias_output = 0 the last layer contributes only x_self_attn, whatever the hotwords are."""
    import torch
    from oracle import paraformer_ref as R
    cfg, W = synth.make_weights(dict(n_enc=1, n_dec=2, contextual=1, vocab=64), seed=5)
    pc = R.PfConfig.from_dict(cfg)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    Wt["decoder.bias_output.weight"] = torch.zeros_like(Wt["decoder.bias_output.weight"])
    torch.manual_seed(0)
    emb, enc = torch.randn(7, 512), torch.randn(19, 512)
    a = R.decoder(emb, enc, 7, Wt, pc, hw_emb=torch.randn(5, 512))
    b = R.decoder(emb, enc, 7, Wt, pc, hw_emb=torch.randn(3, 512) * 9)
    assert torch.equal(a, b)
    Wt["decoder.bias_output.weight"] = torch.from_numpy(W["decoder.bias_output.weight"])
    c = R.decoder(emb, enc, 7, Wt, pc, hw_emb=torch.randn(5, 512))
    assert not torch.equal(a, c)


def test_logprob_topk_oracle_order_and_values():
    import torch
    from oracle import paraformer_ref as R
    x = np.array([[0.0, 2.0, 2.0, -1.0, 1.0, 2.0, 0.5, 0.0]], np.float32)
    lse, lp, ids = R.logprob_topk(x, 4)
    assert list(ids[0]) == [1, 2, 5, 4]                                   # ties by ascending index
    ref = torch.log_softmax(torch.from_numpy(x), -1).numpy()
    assert np.allclose(lp[0], ref[0, ids[0]], atol=1e-6) and abs(lse[0] - np.log(np.exp(x).sum())) < 1e-5
    assert ids[0, 0] == F.find_max(x[0])[1]


def test_cif_matches_closed_form():
    import torch
    from oracle import paraformer_ref as R
    # alphas of exactly 0.5: a token every second frame, each the mean of two frames
    h = torch.arange(8, dtype=torch.float32)[:, None].repeat(1, 4)
    emb, fires = R.cif(h, torch.full((8,), 0.5), 1.0)
    assert emb.shape[0] == 4
    assert torch.allclose(emb[:, 0], torch.tensor([0.5, 2.5, 4.5, 6.5]))
    assert torch.equal(fires >= 1.0, torch.tensor([False, True] * 4))


def test_param_count_is_paraformer_large(synth):
    n = sum(int(np.prod(s)) for s in synth.param_shapes(synth.DEFAULT_CFG).values())
    assert abs(n - 215.8e6) < 0.1e6


def test_flops_formula_matches_survey():
    from oracle import paraformer_ref as R
    assert abs(R.flops(167, 84) / 1e9 - 67.1) < 0.5
    assert abs(R.flops(1000, 501) / 1e9 - 500) < 5


# ---------------------------------------------------------------------------------------------------
# host-side text / timestamp post-processing restatement
# ---------------------------------------------------------------------------------------------------
def _vocab():
    toks = ["<blank>", "<s>", "</s>", "你", "好", "hel@@", "lo", "wor@@", "ld", "a", "b", "<unk>", "ok"]
    return P.Vocab(toks), {t: i for i, t in enumerate(toks)}


def test_cif_and_posenc_oracle_match_reference_compiled_golden():
    """tests/golden/cif_posenc_golden.npz comes from the reference's own compiled ParaformerOnline::CifSearch / GetPosEmb
    (paraformer-online.cpp:240-345) run in offline form (chunk {0,T,0}, last chunk -> tail frame 0.45).  Token counts are
    exact; frames agree to 1e-6 (the online code carries `integrate - threshold` where the offline graph carries
    `alpha - (threshold - integrate)`: equal in exact arithmetic, one rounding apart in fp32)."""
    import torch
    from oracle import paraformer_ref as R
    g = np.load(os.path.join(GOLD, "cif_posenc_golden.npz"))
    for Tn in (1, 7, 33, 167, 1000):
        h, a, ref = g["cif_hidden_%d" % Tn], g["cif_alphas_%d" % Tn], g["cif_frames_%d" % Tn]
        hh = torch.cat([torch.from_numpy(h), torch.zeros(1, h.shape[1])], 0)
        aa = torch.cat([torch.from_numpy(a), torch.tensor([0.45])])
        emb, fires = R.cif(hh, aa, 1.0)
        assert emb.shape[0] == ref.shape[0] == int((fires >= 1.0).sum())
        if ref.size:
            assert np.abs(emb.numpy() - ref).max() <= 1e-6
    pe = R.pos_enc(1000, 560).numpy()
    assert np.abs(pe[:64] - g["pos_emb_64x560"]).max() <= 1e-6
    assert np.abs(pe[999] - g["pos_emb_row1000"]).max() <= 1e-5     # sin/cos of arguments up to 1000 in float


def test_cif_oracle_matches_live_reference_when_built():
    import torch
    from oracle import paraformer_ref as R
    from oracle import text_ref as T
    if not T.available():
        pytest.skip("oracle/_ref/libfunasr_text_ref.so not built (needs /root/reference)")
    rng = np.random.default_rng(12)
    for Tn in (2, 50, 333):
        h = rng.standard_normal((Tn, 16)).astype(np.float32)
        a = rng.uniform(0, 1, Tn).astype(np.float32)
        ref = T.cif_search(h, a)
        emb, _ = R.cif(torch.cat([torch.from_numpy(h), torch.zeros(1, 16)], 0), torch.cat([torch.from_numpy(a), torch.tensor([0.45])]), 1.0)
        assert emb.shape[0] == ref.shape[0] and (ref.size == 0 or np.abs(emb.numpy() - ref).max() <= 1e-6)


def test_text_oracle_matches_reference_compiled_golden():
    """tests/golden/text_golden.json comes from the reference's own compiled vocab.cpp / util.cpp (make_golden.py)."""
    import json
    from oracle import postproc_ref as PP
    g = json.load(open(os.path.join(GOLD, "text_golden.json"), encoding="utf-8"))
    synth = __import__("importlib").import_module("asr-2pass_b200.synth")
    toks = synth.make_tokens(8404)
    cur_lang, v = None, None
    for c in g["text"]:
        if c["lang"] != cur_lang:                       # one stateful vocabulary per language block, as generated
            cur_lang, v = c["lang"], PP.Vocab(toks)
        assert v.vector2string_v2(c["ids"], c["lang"]) == c["text"]
    v = PP.Vocab(toks)
    assert len(g["stamps"]) >= 50
    for c in g["stamps"]:
        assert PP.greedy_search_text(v, c["ids"], "zh-cn", c["us_alphas"], c["us_peaks"]) == c["text"]


def test_text_oracle_matches_live_reference_when_built():
    """In the build container the compiled reference is present: compare on fresh random inputs too."""
    from oracle import postproc_ref as PP
    from oracle import text_ref as T
    if not T.available():
        pytest.skip("oracle/_ref/libfunasr_text_ref.so not built (needs /root/reference)")
    synth = __import__("importlib").import_module("asr-2pass_b200.synth")
    toks = synth.make_tokens(8404)
    rv, ov = T.RefVocab(toks), PP.Vocab(toks)
    rng = np.random.default_rng(99)
    for _ in range(300):
        n = int(rng.integers(0, 30))
        ids = np.where(rng.random(n) < 0.6, rng.integers(3, 7903, n), rng.integers(7903, 8404, n)).astype(np.int32)
        assert rv.vector2string_v2(ids, "zh-cn") == ov.vector2string_v2([int(i) for i in ids], "zh-cn")
    checked = 0
    for _ in range(300):
        n = int(rng.integers(1, 25))
        ids = np.where(rng.random(n) < 0.7, rng.integers(3, 7903, n), rng.integers(7903, 8403, n)).astype(np.int32)
        nf = int(rng.integers(3 * n, 12 * n + 10))
        al = rng.uniform(0, 0.4, nf).astype(np.float32)
        al = (al * ((n + int(rng.integers(-1, 2)) + 1) / al.sum())).astype(np.float32)
        pk = np.zeros(nf, np.float32)
        s = np.float32(0)
        for i in range(nf):
            s = np.float32(s + al[i])
            pk[i] = s
            if s >= np.float32(1 - 1e-4):
                s = np.float32(s - np.float32(1 - 1e-4))
        try:
            want = PP.greedy_search_text(ov, [int(i) for i in ids], "zh-cn", list(al), list(pk))
        except IndexError:
            continue            # undefined in the reference (out-of-bounds read)
        assert rv.greedy_with_stamps(ids, al, pk) == want
        checked += 1
    assert checked > 250


def test_vector2string_v2_rules():
    v, ix = _vocab()
    ids = [ix[t] for t in ["<s>", "你", "好", "hel@@", "lo", "wor@@", "ld", "</s>"]]
    assert v.vector2string_v2(ids, "zh-cn") == "你好hello world"
    v, ix = _vocab()
    assert v.vector2string_v2([ix["a"], ix["b"], ix["ok"]], "zh-cn") == "ab ok"       # single letters are glued
    v, ix = _vocab()
    assert v.vector2string_v2([ix["hel@@"], ix["你"]], "zh-cn") == "hel 你"           # "lo@@ chinese" bad case
    v, ix = _vocab()
    assert v.vector2string_v2([ix["你"], ix["hel@@"]], "zh-cn") == "你hel"           # trailing sub-word: no space


def test_vector2string_v2_carries_state_between_calls():
    v, ix = _vocab()
    assert v.vector2string_v2([ix["ok"]], "zh-cn") == "ok"
    assert v.last_is_complete_english
    assert v.vector2string_v2([ix["ok"]], "zh-cn") == " ok"  # vocab.cpp:177,259-261


def test_timestamp_and_postprocess_plain():
    chars = ["你", "好", "</s>"]
    peaks = np.zeros(30, np.float32)
    peaks[[4, 10, 20]] = 1.0
    alphas = np.full(30, 0.1, np.float32)
    res, ts, cl, _ = P.timestamp_onnx(alphas, peaks, chars)
    assert cl == ["你", "好"]
    rate = np.float32(10.0 * 6 / 1000 / 3)
    assert len(ts) == 2 and ts[0][0] == np.float32(2.5) * rate and ts[0][1] == np.float32(8.5) * rate
    # 30 - 18.5 > 5 -> trailing <sil>, last token ends half way
    assert ts[1][1] == np.float32((30 + 18.5) / 2.0) * rate
    s = P.post_process(["你", "好", "</s>"], ts)
    assert s == "你好 | " + ", ".join(["%f" % float(ts[0][0]), "%f" % float(ts[0][1])]) + "," + \
        ", ".join(["%f" % float(ts[1][0]), "%f" % float(ts[1][1])])


def test_timestamp_rescale_branch_when_peak_count_mismatches():
    chars = ["你", "好", "a"]
    peaks = np.zeros(40, np.float32)
    peaks[[5, 25]] = 1.0                      # 2 peaks but 3 chars + 1 expected -> rescale us_alphas
    alphas = np.full(40, 0.05, np.float32)
    res, ts, cl, al = P.timestamp_onnx(alphas, peaks, chars)
    assert len(ts) == 3
    assert abs(float(np.sum(al)) - 4.0) < 1e-3
    assert all(ts[i][1] <= ts[i + 1][0] + 1e-6 for i in range(2))


def test_timestamp_long_token_is_split_with_sil():
    chars = ["你", "好"]
    peaks = np.zeros(100, np.float32)
    peaks[[3, 50, 97]] = 1.0                   # first token lasts 47 > 30 upsampled frames
    res, ts, _, _ = P.timestamp_onnx(np.full(100, 0.03, np.float32), peaks, chars)
    assert "<sil>" in res and len(ts) == 2
    rate = np.float32(0.02)
    assert abs(float(ts[0][1]) - float((np.float32(1.5) + np.float32(30)) * rate)) < 1e-6


def test_stitch_offline_format():
    text, stamp = P.stitch_offline(["你好 | 0.100000, 0.500000,0.500000, 0.900000", "", "ok | 0.000000, 0.300000"],
                                   [0.0, 1.0, 2.0], "zh-cn")
    assert text == "你好ok"
    assert stamp == "[[100,500],[500,899],[2000,2300]]" or stamp == "[[100,500],[500,900],[2000,2300]]"


def test_fetch_dynamic_rules():
    sr = 16000
    lens = [2 * sr, 3 * sr, 50 * sr, 59 * sr, 61 * sr, 70 * sr]
    b = P.fetch_dynamic(lens, batch_size=8)
    assert b[0] == [0, 1, 2, 3]                # 59 s x 5 would exceed 300 s; the >= 60 s item breaks the batch
    assert b[1] == [4] and b[2] == [5]         # >= 60 s segments go alone
    assert P.fetch_dynamic(lens, 8, use_gpu=False) == [[i] for i in range(6)]  # max_batch = 1 without USE_GPU


# ---------------------------------------------------------------------------------------------------
# host logic of the product: scheduler, model directory, C ABI symbol table
# ---------------------------------------------------------------------------------------------------
def test_scheduler_partitions_cover_everything(synth):
    import importlib
    sch = importlib.import_module("asr-2pass_b200.scheduler")
    lens = synth.segment_lengths(257)
    lens[5] = 100  # too short: zero rows, still scheduled (result "")
    batches = sch.plan_batches(lens, 4096)
    flat = sorted(i for b in batches for i in b)
    assert flat == list(range(257))
    for b in batches:
        assert sum(sch.num_lfr_frames(int(lens[i])) + 1 for i in b if lens[i] >= 400) <= 4096
        assert all(lens[b[k]] <= lens[b[k + 1]] for k in range(len(b) - 1))
    shards = sch.shard_batches(batches, lens, 4)
    assert sorted(i for s in shards for i in s) == list(range(len(batches)))
    for world in (1, 2, 8):
        seen = []
        for r in range(world):
            seen += [i for b in sch.shard_segments(lens, world, r, 4096) for i in b]
        assert sorted(seen) == list(range(257))


def test_frame_count_functions_agree_with_oracle(capi):
    for n in [0, 1, 399, 400, 401, 559, 560, 561, 1359, 1360, 15999, 16000, 159999, 160000, 960000]:
        assert capi.lib().b200pf_num_fbank_frames(n) == F.num_fbank_frames(n)
        assert capi.lib().b200pf_num_lfr_frames(n) == F.num_lfr_frames(F.num_fbank_frames(n))


def test_c_abi_exports_every_declared_symbol(capi):
    hdr = open(os.path.join(ROOT, "include", "b200pf.h")).read()
    declared = sorted(set(re.findall(r"\b(b200pf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    L = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == declared


def test_model_dir_roundtrip_and_probe(capi, synth, modelfile, tmp_path):
    d = str(tmp_path / "m")
    cfg, W, means, vars_, toks = synth.write_synthetic_model_dir(d, dict(n_enc=2, n_dec=1), jitter_ln=True)
    c, n_tok, n_tensors = capi.model_dir_probe(d)
    assert (c.n_enc, c.n_dec, c.vocab, c.feat_dim, c.sample_rate) == (2, 1, 8404, 560, 16000)
    assert n_tok == 8404 and n_tensors == len(W)
    cfg2, W2 = modelfile.read_weights(os.path.join(d, "model.b200pf"))
    assert all(np.array_equal(W[k], W2[k]) for k in W)
    m2, v2 = modelfile.read_am_mvn(os.path.join(d, "am.mvn"))
    assert np.array_equal(m2, means) and np.array_equal(v2, vars_)
    for f in ("am.mvn", "config.yaml", "tokens.json", "model.b200pf"):
        assert os.path.exists(os.path.join(d, f))  # com-define.h:52-88 names + our weight file


def test_model_dir_errors_are_reported(capi, tmp_path):
    with pytest.raises(capi.B200PFError, match="cannot open"):
        capi.model_dir_probe(str(tmp_path / "nope"))


def test_hostile_model_files_become_error_codes(capi, synth, tmp_path):
    """Nothing may unwind through the C ABI: a tensor table whose shape overflows, a table that points past the end of the file,
    and text files the number parsers reject all come back as B200PFError (an error code + text), never as a crash."""
    import shutil
    import struct
    good = str(tmp_path / "good")
    synth.write_synthetic_model_dir(good, dict(n_enc=1, n_dec=1))

    def variant(name):
        d = str(tmp_path / name)
        shutil.copytree(good, d)
        return d

    def weights(recs):
        out = b"B2PFWTS1" + struct.pack("<I", 0) + struct.pack("<I", len(recs))
        for name, dims, offset, nbytes in recs:
            out += name.encode().ljust(96, b"\0") + struct.pack("<I", len(dims)) + struct.pack("<4Q", *(list(dims) + [0] * (4 - len(dims))))
            out += struct.pack("<QQ", offset, nbytes)
        return out

    d = variant("overflow")      # 2^33 x 2^33 elements: the product wraps to 0 in 64 bits once multiplied by 4
    open(os.path.join(d, "model.b200pf"), "wb").write(weights([("encoder.x", (1 << 33, 1 << 33), 0, 0)]))
    with pytest.raises(capi.B200PFError, match="bad shape|size mismatch"):
        capi.model_dir_probe(d)
    d = variant("huge")          # consistent sizes, but far more data than the file holds: no allocation may be attempted
    open(os.path.join(d, "model.b200pf"), "wb").write(weights([("encoder.x", (1 << 20, 1 << 20), 256, 4 << 40)]))
    with pytest.raises(capi.B200PFError, match="bad shape|size mismatch"):
        capi.model_dir_probe(d)
    d = variant("offset")
    open(os.path.join(d, "model.b200pf"), "wb").write(weights([("encoder.x", (4,), 1 << 50, 16)]))
    with pytest.raises(capi.B200PFError, match="size mismatch"):
        capi.model_dir_probe(d)
    d = variant("mvn")           # std::stof would throw on this
    txt = open(os.path.join(d, "am.mvn")).read().split("\n")
    for i, line in enumerate(txt):
        if line.startswith("<LearnRateCoef>"):
            parts = line.split()
            parts[3] = "not-a-number"
            txt[i] = " ".join(parts)
            break
    open(os.path.join(d, "am.mvn"), "w").write("\n".join(txt))
    with pytest.raises(capi.B200PFError):
        capi.model_dir_probe(d)
    d = variant("tokens")        # std::stoul would throw on this escape
    open(os.path.join(d, "tokens.json"), "w").write('["<blank>", "\\uZZZZ"]')
    with pytest.raises(capi.B200PFError):
        capi.model_dir_probe(d)


def test_compute_entry_points_fail_loudly_without_device(capi, synth, tmp_path):
    if capi.device_count() > 0:
        pytest.skip("a B200 is present")
    d = str(tmp_path / "m")
    synth.write_synthetic_model_dir(d, dict(n_enc=1, n_dec=1))
    with pytest.raises(capi.B200PFError, match="no CUDA device"):
        capi.Engine(d)
    with pytest.raises(capi.B200PFError, match="no CUDA device"):
        capi.op_gemm(np.zeros((8, 8), np.float32), np.zeros((8, 8), np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "asr-2pass_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f
