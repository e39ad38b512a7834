"""Generates the committed golden vectors under tests/golden/ IN THE BUILD CONTAINER, where /root/reference
exists: front-end vectors come from the reference's own compiled kaldi-native-fbank (oracle/_ref, built by
oracle/Makefile from the sources where they lie); LFR/CMVN vectors from a literal transcription of
Paraformer::LfrCmvn's vector-insert algorithm (paraformer.cpp:421-461) written independently of the
oracle's index formula; model taps from the fp32 oracle (parity unpinned, see oracle/paraformer_ref.py).

    python tests/golden/make_golden.py
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def lfr_cmvn_literal(feats, means, vars_, lfr_m=7, lfr_n=6):
    """Line-by-line transcription of Paraformer::LfrCmvn (vector inserts and all)."""
    asr_feats = [list(map(np.float32, f)) for f in feats]
    out_feats = []
    T = len(asr_feats)
    T_lrf = int(np.ceil(1.0 * T / lfr_n))
    for _ in range((lfr_m - 1) // 2):
        asr_feats.insert(0, list(asr_feats[0]))
    T = T + (lfr_m - 1) // 2
    for i in range(T_lrf):
        p = []
        if lfr_m <= T - i * lfr_n:
            for j in range(lfr_m):
                p.extend(asr_feats[i * lfr_n + j])
        else:
            num_padding = lfr_m - (T - i * lfr_n)
            for j in range(len(asr_feats) - i * lfr_n):
                p.extend(asr_feats[i * lfr_n + j])
            for _ in range(num_padding):
                p.extend(asr_feats[-1])
        out_feats.append(p)
    out = np.asarray(out_feats, np.float32).reshape(len(out_feats), -1)
    return ((out + means[None, :]).astype(np.float32) * vars_[None, :]).astype(np.float32)


def main():
    from oracle import frontend as F
    synth = importlib.import_module("asr-2pass_b200.synth")
    F.build(ref=True)
    assert F.ref_lib() is not None, "needs /root/reference (run in the build container)"
    means, vars_ = synth.make_cmvn()
    g = {}
    for n in (399, 400, 559, 560, 1359, 16000, 52800):
        pcm16 = synth.make_audio(n, 9000 + n)
        g["pcm_%d" % n] = pcm16
        fb = F.fbank_ref(pcm16.astype(np.float32) / np.float32(32768))
        g["fbank_%d" % n] = fb
    rng = np.random.default_rng(5)
    for n_fb in (1, 3, 6, 7, 12, 13, 98):
        fb = rng.standard_normal((n_fb, 80)).astype(np.float32)
        g["lfr_in_%d" % n_fb] = fb
        g["lfr_out_%d" % n_fb] = lfr_cmvn_literal(fb, means, vars_)
    # knf test-rfft.cc:32-50 known answer through the reference's Rfft
    import ctypes
    d = np.array([1, -1, 3, 8, 20, 6, 0, 2], np.float32)
    F.ref_lib().knf_ref_rfft(d.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 8)
    g["rfft8"] = d
    np.savez_compressed(os.path.join(HERE, "frontend_golden.npz"), **g)

    # small-architecture model taps from the fp32 oracle (2 encoder / 2 decoder layers, jittered LayerNorm)
    import torch
    from oracle import paraformer_ref as R
    cfg, W = synth.make_weights(dict(n_enc=2, n_dec=2), seed=0, jitter_ln=True)
    pc = R.PfConfig.from_dict(cfg)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    m = {}
    for n in (16000, 52800):
        pcm16 = g["pcm_%d" % n]
        feats = F.lfr_cmvn(g["fbank_%d" % n], means, vars_)
        o = R.forward(feats, Wt, pc)
        m["enc_%d" % n] = o["enc"].numpy().astype(np.float16)
        m["alphas_%d" % n] = o["alphas"].numpy()
        m["fires_%d" % n] = o["fires"].numpy()
        m["ids_%d" % n] = np.asarray(o["ids"], np.int32)
        m["token_num_%d" % n] = np.asarray([o["token_num"]], np.int32)
        lg = o["logits"].numpy()
        top2 = np.sort(lg, axis=1)[:, -2:]
        m["top_gap_%d" % n] = (top2[:, 1] - top2[:, 0]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "model_small_golden.npz"), **m)

    # config 3 (timestamp head + contextual decoder + hotword compiler), same small architecture, fp32 oracle
    cfg, W = synth.make_weights(dict(n_enc=2, n_dec=2, timestamp=1, contextual=1), seed=3, jitter_ln=True)
    pc = R.PfConfig.from_dict(cfg)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    rng = np.random.default_rng(77)
    hw_ids = np.zeros((9, 10), np.int32)
    hw_len = np.zeros(9, np.int32)
    for j in range(8):
        L = int(rng.integers(1, 8))
        hw_ids[j, :L] = rng.integers(3, pc.vocab - 1, L)
        hw_len[j] = L
    hw_ids[8, 0], hw_len[8] = 1, 1
    hw = R.select_hotword_rows(R.hotword_embed(hw_ids, Wt), hw_len)
    c3 = dict(hw_ids=hw_ids, hw_len=hw_len, hw_emb=hw.numpy())
    for n in (16000, 52800):
        feats = F.lfr_cmvn(g["fbank_%d" % n], means, vars_)
        o = R.forward(feats, Wt, pc, hw_emb=hw)
        c3["us_alphas_%d" % n] = o["us_alphas"].numpy()
        c3["us_peaks_%d" % n] = o["us_peaks"].numpy()
        c3["ids_%d" % n] = np.asarray(o["ids"], np.int32)
        c3["token_num_%d" % n] = np.asarray([o["token_num"]], np.int32)
        lg = o["logits"].numpy()
        top2 = np.sort(lg, axis=1)[:, -2:]
        c3["top_gap_%d" % n] = (top2[:, 1] - top2[:, 0]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "model_cfg3_golden.npz"), **c3)
    # host text path from the reference's OWN compiled code (oracle/_ref/libfunasr_text_ref.so: Vocab::Vector2StringV2,
    # Vector2String -> TimestampOnnx -> PostProcess)
    import json
    from oracle import text_ref as T
    from oracle import postproc_ref as P
    assert T.available(), "needs oracle/_ref/libfunasr_text_ref.so (make -C oracle ref in the build container)"
    toks = synth.make_tokens(8404)
    rv = T.RefVocab(toks)
    ov = P.Vocab(toks)
    rng = np.random.default_rng(2026)
    text_cases, stamp_cases = [], []
    for lang in ("zh-cn", "en-bpe", ""):
        rvl = T.RefVocab(toks)      # Vector2StringV2 is stateful across calls: one vocabulary per sequence of calls
        for _ in range(40):
            n = int(rng.integers(0, 30))
            ids = np.where(rng.random(n) < 0.6, rng.integers(3, 7903, n), rng.integers(7903, 8404, n))
            ids = np.where(rng.random(n) < 0.06, rng.integers(0, 3, n), ids).astype(np.int32)
            text_cases.append(dict(lang=lang, ids=[int(i) for i in ids], text=rvl.vector2string_v2(ids, lang)))
    while len(stamp_cases) < 60:
        n = int(rng.integers(1, 25))
        ids = np.where(rng.random(n) < 0.7, rng.integers(3, 7903, n), rng.integers(7903, 8403, n)).astype(np.int32)
        nf = int(rng.integers(3 * n, 12 * n + 10))
        al = rng.uniform(0, 0.4, nf).astype(np.float32)
        k = n + (0 if rng.random() < 0.6 else int(rng.integers(-2, 3))) + 1      # != n + 1 exercises the rescale branch
        al = (al * (k / al.sum())).astype(np.float32)
        if rng.random() < 0.15:
            al[: nf // 3] = 0                                                     # long leading silence -> <sil> / token split
        pk = np.zeros(nf, np.float32)
        s = np.float32(0)
        for i in range(nf):
            s = np.float32(s + al[i])
            pk[i] = s
            if s >= np.float32(1 - 1e-4):
                s = np.float32(s - np.float32(1 - 1e-4))
        try:
            P.greedy_search_text(ov, list(ids), "zh-cn", list(al), list(pk))
        except IndexError:
            continue                # the reference reads out of bounds for this input (undefined there): no golden value
        stamp_cases.append(dict(ids=[int(i) for i in ids], us_alphas=[float(x) for x in al], us_peaks=[float(x) for x in pk],
                                text=rv.greedy_with_stamps(ids, al, pk)))
    # CIF integrate-and-fire and the sinusoidal position encoding from the reference's own compiled
    # ParaformerOnline::CifSearch / GetPosEmb (paraformer-online.cpp:240-345)
    cg = {}
    for Tn in (1, 7, 33, 167, 1000):
        h = rng.standard_normal((Tn, 8)).astype(np.float32)
        a = rng.uniform(0, 1, Tn).astype(np.float32)
        if Tn > 20:
            a[10:14] = 0.5                      # exact threshold hits
        cg["cif_hidden_%d" % Tn] = h
        cg["cif_alphas_%d" % Tn] = a
        cg["cif_frames_%d" % Tn] = T.cif_search(h, a)
    cg["pos_emb_64x560"] = T.pos_emb(64, 560)
    cg["pos_emb_row1000"] = T.pos_emb(1000, 560)[999]
    np.savez_compressed(os.path.join(HERE, "cif_posenc_golden.npz"), **cg)
    # VAD segments from the reference's own compiled E2EVadModel (e2e-vad.h), driven in 1 s chunks as Audio::CutSplit does
    vg = {}
    vr = np.random.default_rng(31)
    for k in range(24):
        n = int(vr.integers(1, 6000)) if k % 4 else int(vr.integers(1, 140))
        p = np.full(n, 0.95, np.float32)
        t = int(vr.integers(0, 120))
        while t < n:
            d = int(vr.integers(5, 2500))
            p[t:t + d] = vr.uniform(0.0, 0.08, min(d, n - t)).astype(np.float32)
            t += d + int(vr.integers(5, 300))
        flip = vr.random(n) < float(vr.choice([0.0, 0.02, 0.2]))
        p[flip] = 1 - p[flip]
        if k % 6 == 5:
            p = vr.uniform(0, 1, n).astype(np.float32)
        opts = (int(vr.choice([250, 500, 800])), int(vr.choice([3000, 15000, 20000, 60000])), float(vr.choice([0.6, 0.8, 0.9])))
        vg["p_%d" % k] = p
        vg["opts_%d" % k] = np.asarray(opts, np.float64)
        vg["segs_%d" % k] = T.e2e_vad(p, opts[0], opts[1], opts[2], chunk_frames=100)
    np.savez_compressed(os.path.join(HERE, "vad_segments_golden.npz"), **vg)
    # TimestampSentence (util.cpp:569-637)
    sent_cases = []
    sr = np.random.default_rng(77)
    alphabet = list("一丁七万丈三上下不与") + ["，", "。", "？", "、", ",", "?", "!", ".", " ", "a", "b", "hello", "World", "3", "9", "'", "-", "&", "　", "é", "<unk>"]
    for k in range(300):
        text = "".join(alphabet[int(sr.integers(len(alphabet)))] for _ in range(int(sr.integers(0, 40))))
        t, pairs = 0, []
        for _ in range(int(sr.integers(0, 45))):
            a = t + int(sr.integers(0, 300))
            t = a + int(sr.integers(10, 500))
            pairs.append("[%d,%d]" % (a, t))
        stamp = "" if k % 50 == 0 else ("[]" if k % 77 == 0 else "[" + ",".join(pairs) + "]")
        sent_cases.append(dict(text=text, stamp=stamp, out=T.timestamp_sentence(text, stamp)))
    with open(os.path.join(HERE, "text_golden.json"), "w", encoding="utf-8") as f:
        json.dump(dict(source="reference onnxruntime/src/{vocab,util}.cpp compiled in place (oracle/Makefile ref)", text=text_cases, stamps=stamp_cases,
                       sents=sent_cases),
                  f, ensure_ascii=False)
    make_am_golden(synth)
    make_punc_golden(synth)
    print("wrote", os.listdir(HERE))


def make_punc_golden(synth):
    """Strings from the reference's own compiled CTTransformer::AddPunc + CTokenizer (oracle/am_ref.py) with the scripted network of
    tests/test_punc.py behind the session."""
    import tempfile
    from oracle import am_ref as A
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_punc as TP
    d = tempfile.mkdtemp(prefix="punc_golden_")
    cfg, W, toks = synth.write_synthetic_punc_dir(d, TP.SMALL, seed=0)
    state = dict(seed=0, every=0, seen=[])

    def net(ins):
        state["seen"].append(ins[0][0].tolist())
        cls = TP.scripted_punc(ins[0][0], state["seed"], state["every"]) if state["every"] >= 0 else [1] * ins[0].shape[1]
        lg = np.full((1, len(cls), 6), -5.0, np.float32)
        lg[0, np.arange(len(cls)), cls] = 5.0
        lg[0, :, 5] = 9.0
        return [lg]

    ref = A.RefPunc(d, net, tag="golden")
    cases = []
    for s, text in enumerate(TP._texts(synth, toks, 60)):
        lang = "en-bpe" if s % 5 == 0 else "zh-cn"
        # the reference tokenizer's ids: with a network that never punctuates nothing is ever cut off the cache, so the LAST session
        # call of a request sees every token of the text
        state.update(seed=s, every=-1, seen=[])
        ref.add_punc(text, lang)
        ids = state["seen"][-1] if state["seen"] else []
        for every in (0, 50, 7):
            state.update(seed=s, every=every, seen=[])
            cases.append(dict(text=text, lang=lang, seed=s, every=every, out=ref.add_punc(text, lang), ids=ids))
    ref.close()
    # the realtime model (ct-transformer-online.cpp): multi-turn sessions with the word cache carried from call to call
    ost = dict(seed=0, every=0, vad_pos=0)

    def onet(ins):
        cls = TP._online_script(ins[0][0], ost["seed"], ost["every"], ost["vad_pos"])
        lg = np.full((1, len(cls), 6), -5.0, np.float32)
        lg[0, np.arange(len(cls)), cls] = 5.0
        lg[0, :, 5] = 9.0
        return [lg]

    oref = A.RefPuncOnline(d, onet, tag="golden")
    online = []
    for s, turns in TP._online_sessions(synth, toks, 40):
        cache, rec = [], []
        for text, every in turns:
            ost.update(seed=s, every=every, vad_pos=len(cache))
            out = oref.add_punc(text, cache)
            rec.append(dict(text=text, every=every, out=out, cache=[w.decode("utf-8", "replace") for w in cache]))
        online.append(dict(seed=s, turns=rec))
    oref.close()
    with open(os.path.join(HERE, "punc_golden.json"), "w", encoding="utf-8") as f:
        json.dump(dict(source="reference onnxruntime/src/{ct-transformer,ct-transformer-online,tokenizer}.cpp compiled in place over oracle/fake_ort.cc; scripted network",
                       vocab=TP.SMALL["vocab"], cases=cases, online=online), f, ensure_ascii=False)


def make_am_golden(synth):
    """The reference's own compiled funasr::Paraformer (Forward / CompileHotwordEmbedding) over the stand-in onnxruntime with the
    oracle's network behind the sessions (oracle/am_ref.py): features handed to the session, result strings, hotword id matrices."""
    import tempfile
    import torch
    from oracle import am_ref as A
    from oracle import paraformer_ref as R
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_am_ref_cpu as TA
    assert A.available(), "needs oracle/_ref/libfunasr_am_ref.so (make -C oracle ref in the build container)"
    out = dict(source="reference onnxruntime/src/paraformer.cpp compiled in place over oracle/fake_ort.cc; network = oracle/paraformer_ref.py",
               vocab=8404, seg_dict=TA.SEG_DICT, hotwords=[], forward=[])
    arrays = {}
    for name, extra, seed, hotwords, cases in (
            ("plain", {}, 0, None, ((42, 160000), (3, 16000), (4, 5000), (6, 7120), (11, 48000), (12, 399))),
            ("config3", dict(timestamp=1, contextual=1), 3, "一丁 七万丈 hello", ((42, 160000), (3, 48000), (9, 80000)))):
        d = tempfile.mkdtemp(prefix="am_golden_")
        pc, Wt, means, vars_, toks = TA._model(synth, d, extra, seed)
        sd = os.path.join(d, "seg_dict")
        with open(sd, "w", encoding="utf-8") as f:
            f.write(TA.SEG_DICT)
        seen, hseen = {}, {}
        m = A.RefParaformer(d, TA._net(pc, Wt, seen), n_out=4 if pc.timestamp else 2, hw_net=TA._hw_net(Wt, hseen) if pc.contextual else None,
                            seg_dict=sd, tag="g" + name)
        hw_emb = None
        if pc.contextual:
            # id matrices AND lengths as the reference computes them: a probe network whose output at step t is t + 1 makes the
            # reference's own row selection (paraformer.cpp:678-684) report lengths[j]
            pseen = {}

            def probe(ins):
                pseen["ids"] = ins[0].copy()
                n_hw = ins[0].shape[0]
                return [np.broadcast_to(np.arange(1, 11, dtype=np.float32)[:, None, None], (10, n_hw, pc.d_model)).copy()]
            mp = A.RefParaformer(d, TA._net(pc, Wt, {}), n_out=4, hw_net=probe, seg_dict=sd, tag="probe")
            for hw in TA.HOTWORDS:
                emb = mp.compile_hotwords(hw, dim=pc.d_model)
                out["hotwords"].append(dict(text=hw, ids=pseen["ids"].tolist(), lengths=[int(v) for v in emb[:, 0]]))
            mp.close()
            hw_emb = m.compile_hotwords(hotwords, dim=pc.d_model)
        for k, (aseed, n) in enumerate(cases):
            pcm = synth.make_audio(n, aseed)
            x = pcm.astype(np.float32) / np.float32(32768)
            seen.clear()
            text = m.forward(x, hw_emb)
            case = dict(model=name, cfg=dict(dict(n_enc=2, n_dec=2), **extra), model_seed=seed, audio_seed=aseed, n_samples=n, text=text,
                        hotwords=hotwords)
            if "out" in seen:
                o = seen["out"]
                lp = o["logprobs"].numpy()
                case["ids"] = [int(i) for i in o["ids"]]
                top2 = np.sort(lp[:len(o["ids"])], axis=1)[:, -2:]
                case["gaps"] = [round(float(v), 5) for v in (top2[:, 1] - top2[:, 0])]   # top-1 margin per token (log-prob)
                if n <= 16000:
                    arrays["feats_%s_%d" % (name, k)] = seen["feats"][0]
            out["forward"].append(case)
        m.close()
    # the reference's exported offline API end to end (tests/test_am_ref_cpu.py::test_reference_offline_api_end_to_end)
    capi = importlib.import_module("asr-2pass_b200.capi")
    root = tempfile.mkdtemp(prefix="offline_golden_")
    m = TA._offline_setup(synth, root)
    V, PR = m["V"], m["PR"]

    def vad_net(ins):
        caches = [torch.from_numpy(c[0, :, :, 0].copy()) for c in ins[1:5]]
        sc, nc = V.forward(ins[0][0], m["VWt"], caches)
        return [sc.numpy()[None].astype(np.float32)] + [c.numpy()[None, :, :, None].astype(np.float32) for c in nc]

    def punc_net(ins):
        return [PR.forward(ins[0][0], m["PWt"], m["pcfg"]).numpy()[None].astype(np.float32)]

    out["offline"] = []
    for case in TA.OFFLINE_CASES:
        pcm = TA._offline_audio(synth, case)
        thres, ns, _, _, _ = TA._oracle_offline(capi, m, pcm, case)
        ref = A.RefOffline(m["amd"], TA._net(m["pc"], m["Wt"], {}), am_outputs=4, vad_dir=m["vd"], vad_net=vad_net, vad_thres=thres,
                           punc_dir=m["pd"], punc_net=punc_net)
        text, stamp, sents = ref.infer_buffer(pcm, case["tail"], case["max_len"])
        ref.close()
        out["offline"].append(dict(text=text, stamp=stamp, stamp_sents=sents, n_segments=ns, thres=thres))
    np.savez_compressed(os.path.join(HERE, "am_forward_golden.npz"), **arrays)
    with open(os.path.join(HERE, "am_forward_golden.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False)


if __name__ == "__main__":
    main()
