"""CT-Transformer punctuation.

CPU: the oracle's tokenizer and AddPunc restatement (oracle/punc_ref.py) and the C++ host mirror against the reference's own
compiled tokenizer.cpp / ct-transformer.cpp running over the stand-in onnxruntime (live when oracle/_ref is built; committed
golden strings otherwise).  GPU (-m gpu): the network through the C ABI against the fp32 oracle -- logits within 3e-2 (bf16 GEMM
operands through 4 layers; measured below), classes equal except where the fp32 top-1 margin is below 0.1."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import am_ref as A
from oracle import punc_ref as PR

HERE = os.path.dirname(os.path.abspath(__file__))
SMALL = dict(vocab=6000, d_model=64, n_heads=4, d_ff=128, n_layers=2)


def scripted_punc(ids, seed, period_every):
    """A deterministic stand-in network both sides can evaluate: class from a hash of (token id, position); periods only every
    `period_every`-th token on average (0 = never), so the 200-token cache limit and its forced period are reached."""
    out = []
    for pos, t in enumerate(ids):
        hsh = (int(t) * 2654435761 + pos * 40503 + seed * 97) & 0xFFFFFFFF
        hsh = (hsh >> 7) & 0xFFFF
        if period_every and hsh % period_every == 0:
            out.append(PR.PERIOD if hsh % 3 else PR.QUESTION)
        elif hsh % 11 == 1:
            out.append(PR.COMMA)
        elif hsh % 37 == 2:
            out.append(0)                      # "<unk>" is a class the reference can emit
        else:
            out.append(PR.NOTPUNC)
    return out


def _texts(synth, toks, n):
    out = []
    for s in range(n):
        k = int(np.random.default_rng(1000 + s).integers(0, 900 if s % 7 == 0 else 120))
        t = synth.make_text(k, s, toks) if s % 13 else ""
        if s % 17 == 3:
            t = "  " + t.replace(" ", "  ", 2) + " "
        out.append(t)
    return out


def test_oracle_addpunc_and_tokenizer_match_live_reference(synth, tmp_path):
    if not A.available():
        pytest.skip("oracle/_ref/libfunasr_am_ref.so not built (needs /root/reference)")
    d = str(tmp_path)
    cfg, W, toks = synth.write_synthetic_punc_dir(d, SMALL, seed=0)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    tok = PR.Tokenizer(toks)
    state = dict(mode="net", seed=0, every=0, max_len=0, seen=[])

    def net(ins):
        ids, lens = ins
        assert ids.shape[0] == 1 and int(lens[0]) == ids.shape[1]
        state["max_len"] = max(state["max_len"], ids.shape[1])
        state["seen"].append(ids[0].tolist())
        if state["mode"] == "net":
            return [PR.forward(ids[0], Wt, cfg).numpy()[None].astype(np.float32)]
        cls = scripted_punc(ids[0], state["seed"], state["every"])
        lg = np.full((1, len(cls), 6), -5.0, np.float32)
        lg[0, np.arange(len(cls)), cls] = 5.0
        lg[0, :, 5] = 9.0                      # the sixth class always has the largest logit: the reference must never pick it
        return [lg]

    ref = A.RefPunc(d, net, tag="live")
    texts = _texts(synth, toks, 120)
    for s, text in enumerate(texts):
        lang = "en-bpe" if s % 5 == 0 else "zh-cn"
        state["mode"], state["seen"] = "net", []
        got = ref.add_punc(text, lang)
        mine = []
        exp = PR.add_punc(text, tok, lambda ids: (mine.append(list(ids)), PR.infer_ids(PR.forward(ids, Wt, cfg).numpy()))[1], lang)
        assert got == exp, (s, text[:60])
        assert mine == state["seen"]            # the same id sequences reach the session: tokenizer + cache logic
        for every in (0, 50, 7):
            state.update(mode="script", seed=s, every=every)
            got = ref.add_punc(text, lang)
            exp = PR.add_punc(text, tok, lambda ids: scripted_punc(ids, s, every), lang)
            assert got == exp, (s, every, text[:60])
    assert state["max_len"] > PR.CACHE_POP_TRIGGER_LIMIT     # the forced-period branch was reached
    ref.close()


def test_host_walk_matches_live_reference(capi, synth, tmp_path):
    """funasr_b200::PuncJob / PuncTokenizer (the C++ host mirror) against the reference's compiled AddPunc, same scripted network."""
    if not A.available():
        pytest.skip("oracle/_ref/libfunasr_am_ref.so not built (needs /root/reference)")
    d = str(tmp_path)
    cfg, W, toks = synth.write_synthetic_punc_dir(d, SMALL, seed=0)
    state = dict(seed=0, every=0)

    def net(ins):
        cls = scripted_punc(ins[0][0], state["seed"], state["every"])
        lg = np.full((1, len(cls), 6), -5.0, np.float32)
        lg[0, np.arange(len(cls)), cls] = 5.0
        return [lg]

    ref = A.RefPunc(d, net, tag="hostwalk")
    host = capi.HostPuncTokenizer(toks, synth.PUNC_LIST)
    tok = PR.Tokenizer(toks)
    n = 0
    for s, text in enumerate(_texts(synth, toks, 150)):
        lang = "en-bpe" if s % 5 == 0 else "zh-cn"
        assert host.tokenize(text) == tok.tokenize(text)[1]
        for every in (0, 50, 7):
            state.update(seed=s, every=every)
            assert host.add_punc_scripted(text, lang, s, every) == ref.add_punc(text, lang), (s, every, text[:60])
            n += 1
    assert n == 450
    ref.close()


def test_host_walk_matches_reference_golden(capi, synth):
    g = json.load(open(os.path.join(HERE, "golden", "punc_golden.json"), encoding="utf-8"))
    toks = synth.make_punc_tokens(int(g["vocab"]))
    host = capi.HostPuncTokenizer(toks, synth.PUNC_LIST)
    for c in g["cases"]:
        assert host.tokenize(c["text"]) == c["ids"]
        assert host.add_punc_scripted(c["text"], c["lang"], c["seed"], c["every"]) == c["out"], c["text"][:60]


def _online_script(ids, seed, every, vad_pos):
    return scripted_punc(ids, seed + 13 * vad_pos, every)


def _online_sessions(synth, toks, n_sessions, turns=6):
    for s in range(n_sessions):
        rng = np.random.default_rng(5000 + s)
        out = []
        for turn in range(turns):
            k = int(rng.integers(0, 70 if s % 9 else 300))
            out.append((synth.make_text(k, 100 * s + turn, toks) if rng.random() > 0.1 else "", [0, 50, 7][turn % 3]))
        yield s, out


def test_online_walk_matches_live_reference(capi, synth, tmp_path):
    """The realtime model's AddPunc(text, cache) -- oracle restatement AND C++ host mirror -- against the reference's compiled
    ct-transformer-online.cpp over multi-turn sessions, with a scripted network that depends on the VAD position, and a check that
    the mask the reference hands to its session is VadMask(T, cache size) on both mask inputs."""
    if not A.available():
        pytest.skip("oracle/_ref/libfunasr_am_ref.so not built (needs /root/reference)")
    d = str(tmp_path)
    cfg, W, toks = synth.write_synthetic_punc_dir(d, SMALL, seed=0)
    tok = PR.Tokenizer(toks)
    host = capi.HostPuncTokenizer(toks, synth.PUNC_LIST)
    st = dict(seed=0, every=0, vad_pos=0, bad=0, masked=0)

    def net(ins):
        ids, lens, vm, sm = ins
        T = ids.shape[1]
        assert vm.shape == (1, 1, T, T) and np.array_equal(vm, sm) and int(lens[0]) == T
        if not np.array_equal(vm[0, 0], PR.vad_mask(T, st["vad_pos"])):
            st["bad"] += 1
        st["masked"] += int((vm == 0).any())
        cls = _online_script(ids[0], st["seed"], st["every"], st["vad_pos"])
        lg = np.full((1, T, 6), -5.0, np.float32)
        lg[0, np.arange(T), cls] = 5.0
        lg[0, :, 5] = 9.0
        return [lg]

    ref = A.RefPuncOnline(d, net, tag="live")
    n = 0
    for s, turns in _online_sessions(synth, toks, 80):
        c_ref, c_or, c_host = [], [], []
        for text, every in turns:
            st.update(seed=s, every=every, vad_pos=len(c_ref))
            a = ref.add_punc(text, c_ref)
            b = PR.add_punc_online(text, c_or, tok, lambda ids, v: _online_script(ids, s, every, v))
            c = host.add_punc_online_scripted(text, c_host, s, every)
            assert a == b == c, (s, text[:40])
            assert c_ref == c_or == c_host
            n += 1
    assert n == 480 and st["bad"] == 0 and st["masked"] > 50
    ref.close()


def test_online_walk_matches_reference_golden(capi, synth):
    g = json.load(open(os.path.join(HERE, "golden", "punc_golden.json"), encoding="utf-8"))
    toks = synth.make_punc_tokens(int(g["vocab"]))
    tok = PR.Tokenizer(toks)
    host = capi.HostPuncTokenizer(toks, synth.PUNC_LIST)
    assert len(g["online"]) >= 30
    for sess in g["online"]:
        c_or, c_host = [], []
        for turn in sess["turns"]:
            b = PR.add_punc_online(turn["text"], c_or, tok, lambda ids, v: _online_script(ids, sess["seed"], turn["every"], v))
            c = host.add_punc_online_scripted(turn["text"], c_host, sess["seed"], turn["every"])
            assert b == c == turn["out"], turn["text"][:40]
            assert [w.decode("utf-8", "replace") for w in c_or] == turn["cache"] == [w.decode("utf-8", "replace") for w in c_host]


def test_oracle_addpunc_matches_reference_golden(synth):
    g = json.load(open(os.path.join(HERE, "golden", "punc_golden.json"), encoding="utf-8"))
    toks = synth.make_punc_tokens(int(g["vocab"]))
    tok = PR.Tokenizer(toks)
    assert len(g["cases"]) >= 100
    for c in g["cases"]:
        exp = PR.add_punc(c["text"], tok, lambda ids: scripted_punc(ids, c["seed"], c["every"]), c["lang"])
        assert exp == c["out"], c["text"][:60]
        pieces, ids = tok.tokenize(c["text"])
        assert ids == c["ids"]


def test_punc_graph_restatement_is_self_consistent(synth):
    """forward() is batch independent (a sequence's logits do not depend on others) and emu (bf16 rounding points) stays close."""
    cfg, W = synth.make_punc_weights(SMALL, seed=1)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    ids = np.random.default_rng(0).integers(0, SMALL["vocab"], 57)
    a = PR.forward(ids, Wt, cfg).numpy()
    b = PR.forward(ids, Wt, cfg, emu=True).numpy()
    assert a.shape == (57, 6) and np.isfinite(a).all()
    assert np.abs(a - b).max() < 0.1
    cls = PR.infer_ids(a)
    assert set(cls) <= {0, 1, 2, 3, 4} and len(set(cls)) >= 3


@pytest.mark.gpu
def test_punc_network_against_oracle(capi, synth, gpu, tmp_path):
    for name, over in (("small", SMALL), ("odd", dict(vocab=3000, d_model=172, n_heads=4, d_ff=264, n_layers=2)),
                       ("full", dict(vocab=20000))):
        d = str(tmp_path / name)
        cfg, W, toks = synth.write_synthetic_punc_dir(d, over, seed=2)
        Wt = {k: torch.from_numpy(v) for k, v in W.items()}
        eng = capi.PuncEngine(d, max_tokens=8192)
        assert eng.vocab == cfg["vocab"] and eng.n_punc == 6
        rng = np.random.default_rng(3)
        lens = [1, 2, 15, 16, 17, 20, 33, 64, 200, 431, 5]
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        ids = rng.integers(0, cfg["vocab"], offs[-1]).astype(np.int32)
        cls, lg = eng.infer(ids, offs, logits=True)
        worst, flips = 0.0, 0
        for i, n in enumerate(lens):
            ref = PR.forward(ids[offs[i]:offs[i + 1]], Wt, cfg).numpy()
            mine = lg[offs[i]:offs[i + 1]]
            worst = max(worst, float(np.abs(mine - ref).max()))
            rc = PR.infer_ids(ref)
            top = np.sort(ref[:, :5], axis=1)
            for j in range(n):
                if cls[offs[i] + j] != rc[j]:
                    flips += 1
                    assert top[j, -1] - top[j, -2] < 0.1, (name, i, j)
            assert np.array_equal(cls[offs[i]:offs[i + 1]], PR.infer_ids(mine))     # the device argmax == first max over 5 classes of its own logits
        print("punc %s: max |logit - oracle| = %.4f, class flips %d / %d" % (name, worst, flips, offs[-1]))
        assert worst <= 3e-2
        # batch invariance: a sequence alone gives the same classes and logits
        one, lg1 = eng.infer(ids[offs[8]:offs[9]], np.array([0, lens[8]], np.int32), logits=True)
        assert np.array_equal(one, cls[offs[8]:offs[9]]) and np.array_equal(lg1, lg[offs[8]:offs[9]])
        # empty call, empty sequences inside a call
        e, _ = eng.infer(np.zeros(0, np.int32), np.array([0, 0], np.int32))
        assert len(e) == 0
        c2, _ = eng.infer(ids[:20], np.array([0, 0, 20, 20], np.int32))
        c3, _ = eng.infer(ids[:20], np.array([0, 20], np.int32))
        assert np.array_equal(c2, c3)
        eng.close()


@pytest.mark.gpu
def test_host_addpunc_end_to_end(capi, synth, gpu, tmp_path):
    """CTTransformerB200::AddPunc / AddPuncBatch on the GPU == the oracle's AddPunc walk driven by the GPU's own classes (so that
    bf16 near-ties cannot move the comparison), and AddPuncBatch == one AddPunc per text with fewer engine calls."""
    d = str(tmp_path)
    cfg, W, toks = synth.write_synthetic_punc_dir(d, dict(vocab=20000), seed=5)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    host = capi.HostPunc(d, max_tokens=16384)
    eng = capi.PuncEngine(d, max_tokens=4096)
    tok = PR.Tokenizer(toks)
    texts = [synth.make_text(n, 200 + i, toks) for i, n in enumerate([0, 1, 19, 20, 21, 45, 130, 400, 77, 260, 33, 900])]
    singles, agree, total = [], 0, 0
    for i, text in enumerate(texts):
        lang = "en-bpe" if i % 4 == 3 else "zh-cn"
        got = host.add_punc(text, lang)
        exp = PR.add_punc(text, tok, lambda ids: eng.infer(np.asarray(ids, np.int32), np.array([0, len(ids)], np.int32))[0].tolist(), lang)
        assert got == exp, (i, text[:40])
        fp32 = PR.add_punc(text, tok, lambda ids: PR.infer_ids(PR.forward(ids, Wt, cfg).numpy()), lang)
        agree += got == fp32
        total += 1
        singles.append(got)
    assert agree >= total - 3          # against the fp32 network only a near-tie can change a string
    zh = [t for i, t in enumerate(texts) if i % 4 != 3]
    batch, rounds = host.add_punc_batch(zh, "zh-cn")
    assert batch == [s for i, s in enumerate(singles) if i % 4 != 3]
    assert rounds == max(int(np.ceil(len(tok.tokenize(t)[1]) / 20)) for t in zh)      # lock step: as many rounds as the longest text
    assert host.add_punc("", "zh-cn") == ""
    # continuous batching: concurrent AddPunc calls from several threads join the same lock-step rounds
    import threading
    r0 = host.rounds
    got = [None] * len(zh)

    def work(i):
        got[i] = host.add_punc(zh[i], "zh-cn")

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(zh))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert got == batch
    solo = sum(int(np.ceil(len(tok.tokenize(t)[1]) / 20)) for t in zh)
    assert host.rounds - r0 < solo                      # rounds were shared
    host.close()
    eng.close()


@pytest.mark.gpu
def test_offline_shim_vad_asr_punc_chain(capi, synth, gpu, tmp_path):
    """FunOfflineInit with vad-dir + model-dir + punc-dir: one FunOfflineInferBuffer call runs VAD cut -> batched Paraformer ->
    stitching -> punctuation -> per-sentence stamps, the reference's offline request path (funasrruntime.cpp:208-340)."""
    vd, md, pd = str(tmp_path / "vad"), str(tmp_path / "am"), str(tmp_path / "punc")
    synth.write_synthetic_vad_dir(vd, seed=0)
    cfg, W, means, vars_, am_toks = synth.write_synthetic_model_dir(md, dict(n_enc=2, n_dec=2, timestamp=1), seed=0, jitter_ln=True)
    # the punctuation vocabulary must know the acoustic model's characters: reuse its CJK range (U+4E00...)
    synth.write_synthetic_punc_dir(pd, dict(vocab=20000), seed=5)
    parts = []
    for i, (n_speech, n_sil) in enumerate([(48000, 24000), (80000, 40000), (160000, 32000)]):
        parts += [synth.make_audio(n_speech, 70 + i), np.zeros(n_sil, np.int16)]
    pcm = np.concatenate(parts)
    eng = capi.VadEngine(vd, max_frames=20000)
    p0, _, _, _ = eng.scores(pcm, np.array([0, len(pcm)], np.int64))
    eng.close()
    thres = float(np.clip(1.0 - 2.0 * np.median(p0), 0.05, 0.95))
    base = {"vad-dir": vd, "vad-speech-noise-thres": "%.6f" % thres}
    plain = capi.OfflineHandle(md, max_rows=8192, max_segments=256, batch_size=8, options=base)
    full = capi.OfflineHandle(md, max_rows=8192, max_segments=256, batch_size=8, options=dict(base, **{"punc-dir": pd}))
    punc = capi.HostPunc(pd)
    t0, st0, ss0 = plain.infer_full(pcm, 500, 15000)
    t1, st1, ss1 = full.infer_full(pcm, 500, 15000)
    assert len(t0) > 0 and st0.startswith("[[") and st1 == st0
    assert t1.replace(" ", "") == punc.add_punc(t0, "zh-cn").replace(" ", "") and t1 != t0
    assert ss0 == capi.host_sentence_stamps(t0, st0) and ss1 == capi.host_sentence_stamps(t1, st1)
    assert json.loads(ss1)[0]["ts_list"][0] == json.loads(st1)[0]
    assert sum(len(s["ts_list"]) for s in json.loads(ss1)) <= len(json.loads(st1))
    for h in (plain, full, punc):
        h.close()


@pytest.mark.gpu
def test_realtime_punc_network_and_host(capi, synth, gpu, tmp_path):
    """The realtime model: VadMask attention + causal FSMN (sanm_shift) through the C ABI against the fp32 oracle, and
    CTTransformerOnlineB200::AddPunc over a multi-turn session against the oracle walk driven by the GPU's own classes."""
    d = str(tmp_path)
    cfg, W, toks = synth.write_synthetic_punc_dir(d, dict(vocab=20000, n_layers=3, sanm_shift=5), seed=7)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    eng = capi.PuncEngine(d, max_tokens=4096)
    rng = np.random.default_rng(1)
    lens = [1, 5, 20, 33, 47, 64, 130, 9]
    vps = [0, 3, 7, 32, 1, 70, 66, 8]           # incl. vad_pos <= 1 and >= T: no mask
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ids = rng.integers(0, cfg["vocab"], offs[-1]).astype(np.int32)
    cls, lg = eng.infer(ids, offs, logits=True, vad_pos=vps)
    worst = 0.0
    for i, n in enumerate(lens):
        ref = PR.forward(ids[offs[i]:offs[i + 1]], Wt, cfg, mask=PR.vad_mask(n, vps[i])).numpy()
        worst = max(worst, float(np.abs(lg[offs[i]:offs[i + 1]] - ref).max()))
    print("realtime punc: max |logit - oracle| = %.4f" % worst)
    assert worst <= 3e-2
    cls0, lg0 = eng.infer(ids, offs, logits=True)               # without the mask the masked sequences change
    assert not np.array_equal(lg0[offs[3]:offs[4]], lg[offs[3]:offs[4]]) and np.array_equal(lg0[offs[4]:offs[5]], lg[offs[4]:offs[5]])
    host = capi.HostPuncOnline(d, max_tokens=4096)
    tok = PR.Tokenizer(toks)
    c_host, c_or = [], []
    for turn in range(8):
        text = synth.make_text(int(rng.integers(5, 90)), 900 + turn, toks)
        got = host.add_punc(text, c_host)
        exp = PR.add_punc_online(text, c_or, tok, lambda t, v: eng.infer(np.asarray(t, np.int32), np.array([0, len(t)], np.int32), vad_pos=[v])[0].tolist())
        assert got == exp and c_host == c_or, turn
    host.close()
    eng.close()
