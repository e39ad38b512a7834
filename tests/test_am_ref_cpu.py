"""The reference's own compiled funasr::Paraformer (Forward, CompileHotwordEmbedding) run over a stand-in onnxruntime whose
network is the oracle's restatement of the graph (oracle/am_ref.py, oracle/fake_ort.cc): pins the oracle's and the product's HOST
steps of the hot path -- fbank scale, LFR + CMVN, greedy search over token_num rows, the 4-output timestamp branch, the
result-string protocol, hotword id packing and row selection -- against the reference's code, not against a restatement.
Live tests need oracle/_ref (built from /root/reference in the build container); the golden file they produce
(tests/golden/am_forward_golden.json, make_golden.py) is what the GPU box checks the CUDA path against."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import am_ref as A
from oracle import frontend as F
from oracle import paraformer_ref as R
from oracle import postproc_ref as P

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "am_forward_golden.json")

SEG_DICT = "hello\the@@ llo\nworld\twor@@ ld\nopen\tba daa\nmissing\tzz@@ qqqq\nai\ta i\n"
HOTWORDS = ["", "一丁 七万丈", "hello 一丁", "hello world open", "missing 一丁", "notindict", "一a丁", "一丁" * 7, "㐀一丁",
            "一丁  七", " 一丁", "ai 万 丈三 hello", "A一", "一，丁"]


def needs_ref():
    if not A.available():
        pytest.skip("oracle/_ref/libfunasr_am_ref.so not built (needs /root/reference)")


def _model(synth, tmp, extra=None, seed=0):
    cfg, W, means, vars_, toks = synth.write_synthetic_model_dir(tmp, dict(dict(n_enc=2, n_dec=2), **(extra or {})), seed=seed, jitter_ln=True)
    pc = R.PfConfig.from_dict(cfg)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    return pc, Wt, means, vars_, toks


def _net(pc, Wt, seen):
    def net(ins):
        feats, lens = ins[0], ins[1]
        seen["feats"], seen["lens"] = feats, lens
        hw = torch.from_numpy(ins[2][0]) if len(ins) > 2 else None
        o = R.forward(feats[0], Wt, pc, want_taps=False, hw_emb=hw)
        seen["out"] = o
        outs = [o["logprobs"].numpy()[None].astype(np.float32), np.array([o["token_num"]], np.int64)]
        if pc.timestamp:
            outs += [o["us_alphas"].numpy()[None].astype(np.float32), o["us_peaks"].numpy()[None].astype(np.float32)]
        return outs
    return net


def test_reference_forward_plain_model_pins_frontend_and_text(synth, tmp_path):
    needs_ref()
    pc, Wt, means, vars_, toks = _model(synth, str(tmp_path))
    seen = {}
    m = A.RefParaformer(str(tmp_path), _net(pc, Wt, seen), n_out=2, tag="plain")
    v = P.Vocab(toks)
    for seed, n in ((42, 160000), (3, 16000), (4, 5000), (5, 1359), (6, 960 * 7 + 400)):
        x = synth.make_audio(n, seed).astype(np.float32) / np.float32(32768)
        text = m.forward(x)
        ref_feats = F.lfr_cmvn(F.fbank(x), means, vars_)
        # what the reference hands to its session == the oracle's front end (fbank scale 32768, LFR m=7 n=6, CMVN)
        assert seen["feats"].shape == (1,) + ref_feats.shape and int(seen["lens"][0]) == ref_feats.shape[0]
        assert np.abs(seen["feats"][0] - ref_feats).max() <= 1e-4
        ids = seen["out"]["ids"]
        assert text == v.vector2string_v2(ids, "zh-cn")
    assert m.forward(np.zeros(399, np.float32)) == ""      # shorter than one fbank window (paraformer.cpp:477-480)
    m.close()


def test_reference_forward_timestamp_model_pins_result_string(synth, tmp_path):
    needs_ref()
    pc, Wt, means, vars_, toks = _model(synth, str(tmp_path), dict(timestamp=1))
    seen = {}
    m = A.RefParaformer(str(tmp_path), _net(pc, Wt, seen), n_out=4, tag="ts")
    v = P.Vocab(toks)
    n_stamped = 0
    for seed, n in ((42, 160000), (3, 48000), (9, 80000)):
        x = synth.make_audio(n, seed).astype(np.float32) / np.float32(32768)
        text = m.forward(x)
        o = seen["out"]
        try:
            exp = P.greedy_search_text(v, o["ids"], "zh-cn", o["us_alphas"].numpy(), o["us_peaks"].numpy())
        except IndexError:
            continue          # the reference itself reads out of bounds on such inputs (postproc_ref.py header)
        assert text == exp
        n_stamped += " | " in text
    assert n_stamped >= 2
    m.close()


def _hw_net(Wt, seen):
    def hw_net(ins):
        ids = ins[0]
        seen.setdefault("ids", []).append(ids.copy())
        out = R.hotword_embed(torch.from_numpy(ids.astype(np.int64)), Wt)     # the export's [10, N, D]
        return [out.contiguous().numpy().astype(np.float32)]
    return hw_net


def test_reference_hotword_compile_pins_id_packing(capi, synth, tmp_path):
    needs_ref()
    d = str(tmp_path)
    pc, Wt, means, vars_, toks = _model(synth, d, dict(contextual=1))
    sd = os.path.join(d, "seg_dict")
    with open(sd, "w", encoding="utf-8") as f:
        f.write(SEG_DICT)
    seen = {}
    m = A.RefParaformer(d, _net(pc, Wt, {}), n_out=2, hw_net=_hw_net(Wt, seen), seg_dict=sd, tag="hw")
    for hw in HOTWORDS:
        seen["ids"] = []
        emb = m.compile_hotwords(hw, dim=pc.d_model)
        ref_ids = seen["ids"][-1]
        ids, lens = capi.host_pack_hotwords(toks, hw, sd)
        assert np.array_equal(ids, ref_ids), hw
        # row selection by real length (paraformer.cpp:678-684) == the oracle's select_hotword_rows
        full = R.hotword_embed(torch.from_numpy(ref_ids.astype(np.int64)), Wt)
        exp = R.select_hotword_rows(full, torch.from_numpy(lens.astype(np.int64))).numpy()
        assert emb.shape == exp.shape and np.array_equal(emb, exp.astype(np.float32)), hw
    m.close()


def test_hotword_packing_matches_reference_golden(capi, synth, tmp_path):
    g = json.load(open(GOLD, encoding="utf-8"))
    toks = synth.make_tokens(int(g["vocab"]))
    sd = os.path.join(str(tmp_path), "seg_dict")
    with open(sd, "w", encoding="utf-8") as f:
        f.write(g["seg_dict"])
    assert len(g["hotwords"]) >= 10
    for case in g["hotwords"]:
        ids, lens = capi.host_pack_hotwords(toks, case["text"], sd)
        assert ids.tolist() == case["ids"], case["text"]
        assert lens.tolist() == case["lengths"], case["text"]
