"""CPU suite: the C++ host mirror (libfunasr_b200.so: Detokenizer, TimestampFromPeaks, MergeWithStamps,
StitchSegments, FormBatches) against the literal Python restatement of the reference (oracle/postproc_ref.py)
on seeded random inputs.  String outputs must be byte-identical."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import postproc_ref as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def toks(synth):
    return synth.make_tokens()


def test_host_library_exports_declared_hooks(capi):
    hdr = open(os.path.join(ROOT, "include", "b200pf_host.h")).read()
    declared = sorted(set(re.findall(r"\b(b200pf_host_[a-z0-9_]+)\s*\(", hdr)))
    H = ctypes.CDLL(capi.HOST_LIB_PATH)
    assert not [s for s in declared if not hasattr(H, s)]
    assert sorted(capi.HOST_EXPORTS) == declared
    # the reference's exported C++ symbols (funasrruntime.h:100-138) are present, mangled, in the shim
    import subprocess
    syms = subprocess.run(["nm", "-DC", capi.HOST_LIB_PATH], capture_output=True, text=True).stdout
    for name in ("FunOfflineInit(", "FunOfflineInferBuffer(", "FunOfflineInfer(", "FunOfflineUninit(", "FunASRGetResult(",
                 "FunASRGetStamp(", "FunASRFreeResult(", "FunASRGetRetSnippetTime(", "CompileHotwordEmbedding("):
        assert name in syms, name


def test_detokenizer_matches_oracle_on_random_sequences(capi, toks):
    rng = np.random.default_rng(0)
    d, v = capi.HostDetok(toks), P.Vocab(toks)
    for _ in range(1500):
        n = int(rng.integers(0, 30))
        ids = np.where(rng.random(n) < 0.5, rng.integers(7903, 8404, n), rng.integers(0, 8404, n)).astype(np.int32)
        assert d.text(ids, "zh-cn") == v.vector2string_v2(list(ids), "zh-cn")  # includes the carried-over state


def test_text_for_both_incoming_states_reproduces_the_stateful_detokeniser(capi, toks):
    """What the parallel text assembly relies on (MultiGpuParaformer::AssembleText, ParaformerB200::Decode for large batches):
    a segment's text for an explicit incoming state plus the state it leaves behind, chained in order, is exactly what the
    stateful detokeniser (Vocab::Vector2StringV2 with last_is_complete_english_) produces call after call."""
    rng = np.random.default_rng(3)
    stateful, pure = capi.HostDetok(toks), capi.HostDetok(toks)
    st = False
    flips = 0
    for k in range(1500):
        n = int(rng.integers(0, 30))
        ids = np.where(rng.random(n) < 0.5, rng.integers(7903, 8404, n), rng.integers(0, 8404, n)).astype(np.int32)
        both = [pure.text_state(ids, s, "zh-cn") for s in (False, True)]
        want = stateful.text(ids, "zh-cn")
        assert both[int(st)][0] == want, k
        flips += both[0][0] != both[1][0]
        st = both[int(st)][1]
    assert flips > 50          # the incoming state really matters for many of these sequences


def test_detokenizer_en_bpe(capi):
    toks = ["<blank>", "<s>", "</s>", "▁i", "▁he", "llo", "▁wor", "ld", "'m", "<unk>"]
    d, v = capi.HostDetok(toks), P.Vocab(toks)
    rng = np.random.default_rng(1)
    for _ in range(300):
        ids = rng.integers(0, len(toks), int(rng.integers(0, 12))).astype(np.int32)
        assert d.text(ids, "en-bpe") == v.vector2string_v2(list(ids), "en-bpe")
    assert d.text(np.array([3, 4, 5, 6, 7], np.int32), "en-bpe") == "I hello world"


def test_timestamp_postprocess_matches_oracle(capi, toks):
    rng = np.random.default_rng(2)
    d, v = capi.HostDetok(toks), P.Vocab(toks)
    checked = 0
    for _ in range(600):
        n, Fr = int(rng.integers(1, 20)), int(rng.integers(20, 200))
        ids = np.where(rng.random(n) < 0.4, rng.integers(7903, 8403, n), rng.integers(3, 7903, n)).astype(np.int32)
        peaks = np.zeros(Fr, np.float32)
        npk = n + 1 if rng.random() < 0.6 else int(rng.integers(1, n + 3))  # exercise the rescale branch too
        peaks[np.sort(rng.choice(Fr, size=min(npk, Fr), replace=False))] = 1.0
        alphas = rng.uniform(0, 0.2, Fr).astype(np.float32)
        chars = v.vector2string(list(ids))
        _, ts, _, _ = P.timestamp_onnx(alphas, peaks, chars)
        try:
            want = P.post_process(list(chars), ts)
        except IndexError:   # the reference would read out of bounds here; undefined there, skipped here
            continue
        assert d.timestamp_text(ids, alphas, peaks) == want
        checked += 1
    assert checked > 400


def test_host_text_matches_reference_compiled_golden(capi, toks):
    """The C++ host mirror (csrc/host/text.cpp) against strings produced by the reference's own compiled code."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text_golden.json"), encoding="utf-8"))
    cur_lang, d = None, None
    for c in g["text"]:
        if c["lang"] != cur_lang:
            cur_lang, d = c["lang"], capi.HostDetok(toks)
        assert d.text(np.asarray(c["ids"], np.int32), c["lang"]) == c["text"]
    d = capi.HostDetok(toks)
    for c in g["stamps"]:
        got = d.timestamp_text(np.asarray(c["ids"], np.int32), np.asarray(c["us_alphas"], np.float32), np.asarray(c["us_peaks"], np.float32))
        assert got == c["text"]


def test_stitch_matches_oracle(capi):
    rng = np.random.default_rng(3)
    for _ in range(200):
        n = int(rng.integers(0, 6))
        msgs, starts = [], []
        for _k in range(n):
            r = rng.random()
            if r < 0.2:
                msgs.append("")
            elif r < 0.5:
                msgs.append("text%d" % _k)
            else:
                m = int(rng.integers(1, 4))
                st = ",".join("%f, %f" % (float(np.float32(rng.uniform(0, 5))), float(np.float32(rng.uniform(5, 9)))) for _ in range(m))
                msgs.append("t%d | %s" % (_k, st))
            starts.append(float(np.float32(rng.uniform(0, 100))))
        for lang in ("zh-cn", "en-bpe"):
            assert capi.host_stitch(msgs, starts, lang) == P.stitch_offline(msgs, starts, lang)


def test_sentence_stamps_match_reference_golden(capi):
    """pf::host::SentenceStamps against strings from the reference's compiled TimestampSentence (util.cpp:569-637)."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text_golden.json"), encoding="utf-8"))
    assert len(g["sents"]) >= 200
    for c in g["sents"]:
        assert capi.host_sentence_stamps(c["text"], c["stamp"]) == c["out"], c["text"]


def test_sentence_stamps_match_live_reference_when_built(capi):
    from oracle import text_ref as T
    if not T.available():
        pytest.skip("oracle/_ref/libfunasr_text_ref.so not built (needs /root/reference)")
    rng = np.random.default_rng(0)
    alphabet = list("一丁七万丈三上下不与") + ["，", "。", "？", "、", ",", "?", "!", ".", " ", "a", "b", "hello", "World", "3", "9", "'", "-", "&", "　", "é", "ü"]
    for k in range(2000):
        text = "".join(alphabet[int(rng.integers(len(alphabet)))] for _ in range(int(rng.integers(0, 40))))
        t, pairs = 0, []
        for _ in range(int(rng.integers(0, 45))):
            a = t + int(rng.integers(0, 300))
            t = a + int(rng.integers(10, 500))
            pairs.append("[%d,%d]" % (a, t))
        stamp = "" if k % 50 == 0 else ("[]" if k % 77 == 0 else "[" + ",".join(pairs) + "]")
        assert capi.host_sentence_stamps(text, stamp) == T.timestamp_sentence(text, stamp), (text, stamp)


def test_checkpoint_converter_round_trip(synth, tmp_path):
    """tools/convert_funasr.py on synthetic state_dicts (no real checkpoint exists offline): dimensions are recovered from the tensor
    shapes, flags from the presence of the config-3 tensors, and the written file holds the same bytes."""
    import importlib.util
    import json
    import torch
    mf = importlib.import_module("asr-2pass_b200.modelfile")
    spec = importlib.util.spec_from_file_location("convert_funasr", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "convert_funasr.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    d = str(tmp_path)
    cfg, W = synth.make_weights(dict(n_enc=3, n_dec=2, timestamp=1, contextual=1, d_model=64, d_ff=128, feat_dim=80, vocab=300, n_heads=4), seed=1)
    sd = {k: torch.from_numpy(v) for k, v in W.items()}
    sd["some.unused.buffer"] = torch.zeros(3)
    torch.save({"state_dict": sd}, os.path.join(d, "am.pt"))
    with open(os.path.join(d, "am.yaml"), "w") as f:
        f.write("encoder_conf:\n  attention_heads: 4\npredictor_conf:\n  threshold: 1.0\n  tail_threshold: 0.45\n")
    conv.main(["am", "--checkpoint", os.path.join(d, "am.pt"), "--config", os.path.join(d, "am.yaml"), "--out-dir", os.path.join(d, "am")])
    c2, W2 = mf.read_weights(os.path.join(d, "am", "model.b200pf"))
    for k in ("feat_dim", "d_model", "n_heads", "d_ff", "n_enc", "n_dec", "kernel", "vocab", "timestamp", "contextual"):
        assert int(c2[k]) == int(cfg[k]), k
    assert set(W2) == set(W) and all(np.array_equal(W2[k], W[k]) for k in W)
    # hyper-parameters the kernels hard-code: a config.yaml that carries another value stops the conversion instead of
    # producing a model file that decodes wrong text
    for bad in ("predictor_conf:\n  smooth_factor: 0.8\n", "predictor_conf:\n  noise_threshold: 0.01\n", "encoder_conf:\n  sanm_shfit: 3\n",
                "predictor_conf:\n  use_cif1_cnn: true\n", "predictor_conf:\n  upsample_type: cnn\n", "decoder_conf:\n  sanm_shfit: 5\n",
                "frontend_conf:\n  lfr_m: 5\n"):
        with open(os.path.join(d, "bad.yaml"), "w") as f:
            f.write(bad)
        with pytest.raises(SystemExit, match="unsupported"):
            conv.main(["am", "--checkpoint", os.path.join(d, "am.pt"), "--config", os.path.join(d, "bad.yaml"), "--out-dir", os.path.join(d, "bad")])
    with open(os.path.join(d, "ok.yaml"), "w") as f:      # the supported values, spelled out, pass
        f.write("encoder_conf:\n  sanm_shfit: 0\n  attention_heads: 4\npredictor_conf:\n  smooth_factor: 1.0\n  noise_threshold: 0\n  use_cif1_cnn: false\n"
                "  upsample_type: cnn_blstm\n  upsample_times: 3\nfrontend_conf:\n  lfr_m: 7\n  lfr_n: 6\n")
    conv.main(["am", "--checkpoint", os.path.join(d, "am.pt"), "--config", os.path.join(d, "ok.yaml"), "--out-dir", os.path.join(d, "ok")])
    pcfg, PW = synth.make_punc_weights(dict(vocab=500, d_model=64, n_heads=4, d_ff=128, n_layers=3, sanm_shift=5), seed=2)
    torch.save({k: torch.from_numpy(v) for k, v in PW.items()}, os.path.join(d, "punc.pt"))
    with open(os.path.join(d, "punc.yaml"), "w", encoding="utf-8") as f:
        f.write("encoder_conf:\n  attention_heads: 4\n  sanm_shfit: 5\nmodel_conf:\n  punc_list:\n" + "".join('  - "%s"\n' % p for p in synth.PUNC_LIST))
    conv.main(["punc", "--checkpoint", os.path.join(d, "punc.pt"), "--config", os.path.join(d, "punc.yaml"), "--out-dir", os.path.join(d, "punc")])
    c3, W3 = mf.read_weights(os.path.join(d, "punc", "punc.b200pf"))
    assert {k: int(c3[k]) for k in ("vocab", "d_model", "n_heads", "d_ff", "n_layers", "kernel", "n_punc", "sanm_shift")} == \
        {k: int(pcfg[k]) for k in ("vocab", "d_model", "n_heads", "d_ff", "n_layers", "kernel", "n_punc", "sanm_shift")}
    assert json.load(open(os.path.join(d, "punc", "punc_list.json"), encoding="utf-8")) == synth.PUNC_LIST
    assert all(np.array_equal(W3[k], PW[k]) for k in PW)
    VW = synth.make_vad_weights(3)
    torch.save({k: torch.from_numpy(v) for k, v in VW.items()}, os.path.join(d, "vad.pt"))
    conv.main(["vad", "--checkpoint", os.path.join(d, "vad.pt"), "--out-dir", os.path.join(d, "vad")])
    _, W4 = mf.read_weights(os.path.join(d, "vad", "vad.b200pf"))
    assert all(np.array_equal(W4[k], VW[k]) for k in VW)


def test_pruned_posteriors_expand_to_the_rows_the_lm_decoders_read(capi):
    """SURVEY.md §8(f) rank 3 host side: pf::host::ExpandPrunedPosteriors rebuilds the dense log-softmax rows WfstDecoder::Search
    (wfst-decoder.cpp:27-57) / CtcPrefixDecoder::CtcSearch (ctc-prefix-decoder.cpp:157) read from the engine's top-k output: the k
    listed classes keep their exact log-probabilities, every row stays a normalised distribution, no unlisted class outranks a
    listed one, and a top-10 prune of the rebuilt row (CtcPrefixDecoder's first_beam_size) equals the top-10 of the full row."""
    from oracle import paraformer_ref as R
    rng = np.random.default_rng(5)
    V, k = 8404, 16
    logits = (rng.standard_normal((23, V)) * 2.5).astype(np.float32)
    logits[4] *= 6.0                                  # a peaky row: the listed classes hold nearly all the mass
    lse, lp, ids = R.logprob_topk(logits, k)
    full = logits - lse[:, None]
    dense = capi.host_expand_posteriors(lp, ids, V)
    assert dense.shape == (23, V) and np.isfinite(dense).all()
    for r in range(23):
        assert np.array_equal(dense[r, ids[r]], lp[r].astype(np.float32))                     # listed classes: exact
        assert abs(float(np.exp(dense[r].astype(np.float64)).sum()) - 1.0) <= 1e-4             # still a distribution
        rest = np.delete(dense[r], ids[r])
        assert np.all(rest == rest[0]) and rest[0] <= lp[r].min() + 1e-6                      # the rest: one floor value below the list
        top_full = np.lexsort((np.arange(V), -full[r]))[:10]
        top_dense = np.lexsort((np.arange(V), -dense[r]))[:10]
        assert np.array_equal(top_full, top_dense)
