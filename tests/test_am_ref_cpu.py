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


# ---- the reference's exported offline API end to end (funasrruntime.cpp FunOfflineInit / FunOfflineInferBuffer) ---------------
OFFLINE_PUNC = dict(vocab=20000, d_model=64, n_heads=4, d_ff=128, n_layers=2)
OFFLINE_CASES = [dict(bursts=[(48000, 24000), (80000, 40000), (160000, 32000)], seed=70, tail=500, max_len=15000, q=0.5),
                 dict(bursts=[(30000, 9000), (12000, 30000), (90000, 5000), (20000, 16000)], seed=90, tail=250, max_len=3000, q=0.4),
                 dict(bursts=[(8000, 0)], seed=95, tail=800, max_len=60000, q=0.5)]


def _offline_setup(synth, root):
    from oracle import punc_ref as PR
    from oracle import vad_ref as V
    amd, vd, pd = (os.path.join(root, x) for x in ("am", "vad", "punc"))
    for x in (amd, vd, pd):
        os.makedirs(x)
    pc, Wt, means, vars_, toks = _model(synth, amd, dict(timestamp=1))
    VW, vmeans, vvars = synth.write_synthetic_vad_dir(vd, seed=0)
    VWt = {k: torch.from_numpy(v) for k, v in VW.items()}
    pcfg, PW, ptoks = synth.write_synthetic_punc_dir(pd, OFFLINE_PUNC, seed=5)
    PWt = {k: torch.from_numpy(v) for k, v in PW.items()}
    return dict(amd=amd, vd=vd, pd=pd, pc=pc, Wt=Wt, means=means, vars=vars_, toks=toks, VWt=VWt, vmeans=vmeans, vvars=vvars, pcfg=pcfg,
                PWt=PWt, ptoks=ptoks, PR=PR, V=V)


def _offline_audio(synth, case):
    parts = []
    for i, (ns, nsil) in enumerate(case["bursts"]):
        parts += [synth.make_audio(ns, case["seed"] + i), np.zeros(nsil, np.int16)]
    return np.concatenate(parts)


def _oracle_offline(capi, m, pcm, case):
    """The oracle's composition of the offline request path: whole-recording VAD scores -> SegmentVad -> per segment fbank / LFR /
    network / greedy search with stamps (in ascending-length order, like the reference calls Forward) -> stitching -> AddPunc ->
    sentence stamps."""
    PR, V = m["PR"], m["V"]
    x = pcm.astype(np.float32) / np.float32(32768)
    p0 = V.forward(V.lfr_cmvn(F.fbank(x), m["vmeans"], m["vvars"]), m["VWt"])[0].numpy()[:, 0]
    thres = float(np.clip(1.0 - 2.0 * np.quantile(p0, case["q"]), 0.05, 0.95))
    segs = capi.host_vad_segments(p0, case["tail"], case["max_len"], thres)
    v = P.Vocab(m["toks"])
    msgs = [None] * len(segs)
    for i in sorted(range(len(segs)), key=lambda k: int(segs[k][1]) - int(segs[k][0])):
        seg = x[int(segs[i][0]) * 16:int(segs[i][1]) * 16]
        o = R.forward(F.lfr_cmvn(F.fbank(seg), m["means"], m["vars"]), m["Wt"], m["pc"], want_taps=False)
        msgs[i] = P.greedy_search_text(v, o["ids"], "zh-cn", o["us_alphas"].numpy(), o["us_peaks"].numpy())
    text, stamp = P.stitch_offline(msgs, [float(int(s[0]) * 16) / 16000 for s in segs], "zh-cn")
    tok = PR.Tokenizer(m["ptoks"])
    text = PR.add_punc(text, tok, lambda ids: PR.infer_ids(PR.forward(ids, m["PWt"], m["pcfg"]).numpy()), "zh-cn")
    sents = capi.host_sentence_stamps(text, stamp) if stamp else ""
    return thres, len(segs), text, stamp, sents


def test_reference_offline_api_end_to_end(capi, synth, tmp_path):
    """The reference's OWN FunOfflineInit / FunOfflineInferBuffer (funasrruntime.cpp, offline-stream.cpp, audio.cpp, fsmn-vad*.cpp,
    paraformer.cpp, ct-transformer.cpp, util.cpp compiled in place) on model directories whose sessions are served by the oracle's
    networks: LoadPcmwav, CutSplit, length sort + un-permute, Forward per segment, stitching, punctuation, TimestampSentence.  Its
    text / stamp / stamp_sents must equal the oracle's composition of the same path."""
    needs_ref()
    m = _offline_setup(synth, str(tmp_path))
    V, PR = m["V"], m["PR"]

    def vad_net(ins):
        caches = [torch.from_numpy(c[0, :, :, 0].copy()) for c in ins[1:5]]
        sc, nc = V.forward(ins[0][0], m["VWt"], caches)
        return [sc.numpy()[None].astype(np.float32)] + [c.numpy()[None, :, :, None].astype(np.float32) for c in nc]

    def punc_net(ins):
        return [PR.forward(ins[0][0], m["PWt"], m["pcfg"]).numpy()[None].astype(np.float32)]

    n_seg = 0
    for k, case in enumerate(OFFLINE_CASES):
        pcm = _offline_audio(synth, case)
        thres, ns, text, stamp, sents = _oracle_offline(capi, m, pcm, case)
        # a fresh handle per case: the threshold lives in the VAD directory's config.yaml
        ref = A.RefOffline(m["amd"], _net(m["pc"], m["Wt"], {}), am_outputs=4, vad_dir=m["vd"], vad_net=vad_net, vad_thres=thres,
                           punc_dir=m["pd"], punc_net=punc_net)
        got = ref.infer_buffer(pcm, case["tail"], case["max_len"])
        ref.close()
        assert got == (text, stamp, sents), k
        n_seg += ns
    assert n_seg >= 5


def test_oracle_offline_flow_matches_reference_golden(capi, synth, tmp_path):
    """Same comparison against the strings the reference's API produced in the build container (am_forward_golden.json)."""
    g = json.load(open(GOLD, encoding="utf-8"))
    m = _offline_setup(synth, str(tmp_path))
    assert len(g["offline"]) == len(OFFLINE_CASES)
    for case, gold in zip(OFFLINE_CASES, g["offline"]):
        _, ns, text, stamp, sents = _oracle_offline(capi, m, _offline_audio(synth, case), case)
        assert (text, stamp, sents) == (gold["text"], gold["stamp"], gold["stamp_sents"]) and ns == gold["n_segments"]
