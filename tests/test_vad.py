"""FSMN-VAD scores: oracle consistency on the CPU, GPU parity through the C ABI (-m gpu).

Tolerances: LFR + CMVN features <= 1e-4 (measured 1e-6; fbank in fp32 + a double FFT, as for the acoustic model); silence
probability and the 248-way softmax <= 5e-2 absolute, log-probabilities <= 0.3, against the fp32 oracle (measured 4e-2 /
0.22: bf16 GEMM operands through ten layers of a synthetic model whose silence logit was made deliberately large and
sensitive; the consumer compares 1 - p_sil against speech_noise_thres = 0.6-0.9, e2e-vad.h:602-640)."""
import importlib

import numpy as np
import pytest

from oracle import frontend as F
from oracle import vad_ref as V


def test_vad_oracle_chunked_equals_whole(synth):
    """The reference feeds 1 s chunks and carries four [128 x 19] caches (fsmn-vad.cpp:96-133); with a left-context-only
    memory block that equals one pass over the whole recording, which is what the GPU path computes."""
    import torch
    W = synth.make_vad_weights(0)
    x = np.random.default_rng(0).standard_normal((237, 400)).astype(np.float32)
    full, _ = V.forward(x, W)
    caches, parts = None, []
    for a in range(0, 237, 100):
        s, caches = V.forward(x[a:a + 100], W, caches)
        assert all(c.shape == (128, 19) for c in caches)
        parts.append(s)
    assert torch.equal(torch.cat(parts), full)
    assert torch.allclose(full.sum(-1), torch.ones(237), atol=1e-5)
    assert float(full.std()) > 1e-3          # the synthetic weights give a non-trivial posterior


def test_vad_lfr_matches_literal_transcription():
    """FsmnVad::LfrCmvn (fsmn-vad.cpp:182-224) transcribed with its vector inserts vs the oracle's index formula."""
    rng = np.random.default_rng(1)
    for T in (1, 2, 3, 5, 17):
        fb = rng.standard_normal((T, 80)).astype(np.float32)
        means, vars_ = rng.standard_normal(400).astype(np.float32), rng.uniform(0.5, 2, 400).astype(np.float32)
        feats = [list(f) for f in fb]
        m, n = 5, 1
        T_lfr = int(np.ceil(T / n))
        for _ in range((m - 1) // 2):
            feats.insert(0, list(feats[0]))
        Tp = T + (m - 1) // 2
        out = []
        for i in range(T_lfr):
            p = []
            if m <= Tp - i * n:
                for j in range(m):
                    p.extend(feats[i * n + j])
            else:
                for j in range(len(feats) - i * n):
                    p.extend(feats[i * n + j])
                for _ in range(m - (Tp - i * n)):
                    p.extend(feats[-1])
            out.append(p)
        lit = ((np.asarray(out, np.float32) + means[None]) * vars_[None]).astype(np.float32)
        assert np.array_equal(V.lfr_cmvn(fb, means, vars_), lit)


def test_vad_segmenter_matches_reference_compiled_golden(capi):
    """pf::host::SegmentVad against segments produced by the reference's own compiled E2EVadModel (tests/golden/
    vad_segments_golden.npz, make_golden.py), driven in 1 s chunks the way Audio::CutSplit drives it."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vad_segments_golden.npz"))
    total = 0
    for k in range(24):
        mes, msl, th = g["opts_%d" % k]
        got = capi.host_vad_segments(g["p_%d" % k], int(mes), int(msl), float(th))
        assert np.array_equal(got, g["segs_%d" % k]), k
        total += len(got)
    assert total > 50


def test_streaming_vad_equals_whole_recording_for_any_chunking(capi):
    """pf::host::StreamingVad (the 2-pass stream's form: scores pushed chunk by chunk, segments reported as they close) returns,
    in total, exactly the segments the reference's compiled E2EVadModel produced for the same scores -- for chunks of 1 frame,
    60 frames (the 2-pass server's 600 ms), 100 frames and random sizes."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vad_segments_golden.npz"))
    rng = np.random.default_rng(3)
    for k in range(24):
        mes, msl, th = g["opts_%d" % k]
        p = g["p_%d" % k]
        ref = [tuple(int(v) for v in r) for r in g["segs_%d" % k]]
        for chunks in ([1] * len(p), [60] * (len(p) // 60 + 1), [100] * (len(p) // 100 + 1), list(rng.integers(1, 300, len(p) // 20 + 2))):
            c, tot = [], 0
            for n in chunks:                       # trim the list to cover exactly len(p) frames
                if tot >= len(p):
                    break
                c.append(int(min(n, len(p) - tot)))
                tot += c[-1]
            assert capi.host_vad_segments_streaming(p, c, int(mes), int(msl), float(th)) == ref, (k, c[:4])


def test_vad_segmenter_matches_live_reference_when_built(capi):
    from oracle import text_ref as T
    if not T.available():
        pytest.skip("oracle/_ref/libfunasr_text_ref.so not built (needs /root/reference)")
    rng = np.random.default_rng(5)
    segs = 0
    for k in range(400):
        n = int(rng.integers(1, 5000)) if k % 3 else int(rng.integers(1, 130))
        p = np.full(n, 0.95, np.float32)
        t = int(rng.integers(0, 120))
        while t < n:
            d = int(rng.integers(5, 2500))
            p[t:t + d] = rng.uniform(0.0, 0.08, min(d, n - t)).astype(np.float32)
            t += d + int(rng.integers(5, 300))
        flip = rng.random(n) < float(rng.choice([0.0, 0.02, 0.2, 0.4]))
        p[flip] = 1 - p[flip]
        if k % 7 == 0:
            p = rng.uniform(0, 1, n).astype(np.float32)
        if k % 17 == 0:
            p = np.where(rng.random(n) < 0.5, np.float32(0.1), np.float32(0.1000001)).astype(np.float32)   # on the threshold
        mes, msl, th = int(rng.choice([250, 500, 800])), int(rng.choice([3000, 15000, 60000])), float(rng.choice([0.6, 0.8, 0.9]))
        ref = T.e2e_vad(p, mes, msl, th, chunk_frames=int(rng.choice([100, 98, 37])))
        assert np.array_equal(capi.host_vad_segments(p, mes, msl, th), ref)
        segs += len(ref)
    assert segs > 1000


@pytest.mark.gpu
def test_vad_scores_against_oracle(capi, synth, gpu, tmp_path):
    d = str(tmp_path)
    W, means, vars_ = synth.write_synthetic_vad_dir(d, seed=0)
    eng = capi.VadEngine(d, max_frames=20000)
    lens = [16000, 400, 399, 52800, 160000, 0, 1359]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pcm = np.concatenate([synth.make_audio(n, 40 + i) if n else np.zeros(0, np.int16) for i, n in enumerate(lens)])
    p0, fo, probs, feats = eng.scores(pcm, offs, all_probs=True, feats=True)
    assert list(np.diff(fo)) == [F.num_fbank_frames(n) for n in lens]
    for i, n in enumerate(lens):
        T = F.num_fbank_frames(n)
        if T == 0:
            continue
        x = pcm[offs[i]:offs[i + 1]].astype(np.float32) / np.float32(32768)
        ref_feats = V.lfr_cmvn(F.fbank(x), means, vars_)
        assert np.abs(feats[fo[i]:fo[i + 1]] - ref_feats).max() <= 1e-4
        ref, _ = V.forward(ref_feats, W)
        ref = ref.numpy()
        assert np.abs(probs[fo[i]:fo[i + 1]] - ref).max() <= 5e-2
        assert np.abs(p0[fo[i]:fo[i + 1]] - ref[:, 0]).max() <= 5e-2
        assert np.abs(np.log(probs[fo[i]:fo[i + 1]] + 1e-30) - np.log(ref + 1e-30)).max() <= 0.3
        assert np.allclose(probs[fo[i]:fo[i + 1]].sum(1), 1.0, atol=1e-4)
        assert np.array_equal(p0[fo[i]:fo[i + 1]], probs[fo[i]:fo[i + 1], 0])
    # batch invariance: a recording's scores do not depend on what it is processed with
    one, fo1, _, _ = eng.scores(pcm[offs[3]:offs[4]], np.array([0, lens[3]], np.int64))
    assert np.array_equal(one, p0[fo[3]:fo[4]])
    eng.close()


@pytest.mark.gpu
def test_vad_one_hour_stream(capi, synth, gpu, tmp_path):
    """Throughput sanity at the size the survey names (a 1 h stream): scores for 360 k frames in one call."""
    import time
    d = str(tmp_path)
    synth.write_synthetic_vad_dir(d, seed=0)
    eng = capi.VadEngine(d, max_frames=400000)
    blk = synth.make_audio(16000 * 60, 9)
    pcm = np.tile(blk, 60)
    offs = np.array([0, len(pcm)], np.int64)
    eng.scores(pcm[:16000 * 60], np.array([0, 16000 * 60], np.int64))
    t0 = time.perf_counter()
    p0, fo, _, _ = eng.scores(pcm, offs)
    dt = time.perf_counter() - t0
    assert len(p0) == 1 + (len(pcm) - 400) // 160 and np.isfinite(p0).all()
    assert np.array_equal(p0[100:5000], p0[6000 + 100:6000 + 5000])       # the tiled minute repeats once the 19-frame history is the same
    print("VAD 1 h stream: %.1f ms -> %.0f x real time" % (dt * 1e3, 3600.0 / dt))
    assert 3600.0 / dt > 1000
    eng.close()


@pytest.mark.gpu
def test_offline_shim_vad_cut_path(capi, synth, gpu, tmp_path):
    """FunOfflineInit with "vad-dir": FunOfflineInferBuffer cuts the recording with GPU scores + the E2E state machine
    (the reference's UseVad() branch, funasrruntime.cpp:243-245 -> Audio::CutSplit) and decodes exactly those segments."""
    vd, md = str(tmp_path / "vad"), str(tmp_path / "am")
    import os
    os.makedirs(vd), os.makedirs(md)
    synth.write_synthetic_vad_dir(vd, seed=0)
    synth.write_synthetic_model_dir(md, dict(n_enc=2, n_dec=2), seed=0, jitter_ln=True)
    # bursts of synthetic speech separated by digital silence, 34 s
    parts = []
    for i, (n_speech, n_sil) in enumerate([(48000, 24000), (80000, 40000), (16000, 16000), (160000, 32000), (64000, 44000)]):
        parts += [synth.make_audio(n_speech, 70 + i), np.zeros(n_sil, np.int16)]
    pcm = np.concatenate(parts)
    eng = capi.VadEngine(vd, max_frames=20000)
    p0, fo, _, _ = eng.scores(pcm, np.array([0, len(pcm)], np.int64))
    eng.close()
    # synthetic weights: put the speech/noise threshold where the scores actually are (speech <=> p_sil <= (1 - thres) / 2)
    thres = float(np.clip(1.0 - 2.0 * np.median(p0), 0.05, 0.95))
    h = capi.OfflineHandle(md, max_rows=8192, max_segments=256, batch_size=8, vad_dir=vd, vad_thres=thres)
    for tail, mx in ((800, 15000), (250, 3000)):
        cut = h.vad_cut(pcm, tail, mx)
        assert np.array_equal(cut, capi.host_vad_segments(p0, tail, mx, thres))
        assert len(cut) >= 1 and (cut[:, 1] > cut[:, 0]).all() and (np.diff(cut[:, 0]) > 0).all()
        assert (cut[:, 1] - cut[:, 0]).max() <= mx + 1000
        text, stamp = h.infer_buffer_vad(pcm, tail, mx)
        b = np.minimum(cut[:, 0].astype(np.int64) * 16, len(pcm))
        e = np.minimum(cut[:, 1].astype(np.int64) * 16, len(pcm))
        keep = e > b
        expect = h.infer_segments(pcm, b[keep], e[keep])
        expect = expect[0] if isinstance(expect, tuple) else expect
        assert text.replace(" ", "") == expect.replace(" ", "")
    # all-silence recording: no segments, empty text, still a result
    sil = np.zeros(48000, np.int16)
    if len(h.vad_cut(sil, 800, 15000)) == 0:
        assert h.infer_buffer_vad(sil, 800, 15000)[0] == ""
    h.close()


@pytest.mark.gpu
def test_tpass_offline_leg_equals_offline_api_on_the_same_stream(capi, synth, gpu, tmp_path):
    """FunTpassInit / FunTpassOnlineInit / FunTpassInferBuffer (funasrruntime.h:121-132), the 2-pass server's entry points: a stream
    fed in 600 ms chunks closes the SAME VAD segments as the whole recording through FunOfflineInferBuffer, and every closed
    segment's corrected text (`tpass_msg`, mode "2pass-offline") is the offline model's text for that segment.  The streaming
    model's partial results are outside this path: `msg` stays empty."""
    import os
    vd, md = str(tmp_path / "vad"), str(tmp_path / "am")
    os.makedirs(vd), os.makedirs(md)
    synth.write_synthetic_vad_dir(vd, seed=0)
    synth.write_synthetic_model_dir(md, dict(n_enc=2, n_dec=2), seed=0, jitter_ln=True)
    parts = []
    for i, (n_speech, n_sil) in enumerate([(48000, 24000), (80000, 40000), (16000, 16000), (160000, 32000), (64000, 44000)]):
        parts += [synth.make_audio(n_speech, 70 + i), np.zeros(n_sil, np.int16)]
    pcm = np.concatenate(parts)
    eng = capi.VadEngine(vd, max_frames=20000)
    p0, _, _, _ = eng.scores(pcm, np.array([0, len(pcm)], np.int64))
    eng.close()
    thres = float(np.clip(1.0 - 2.0 * np.median(p0), 0.05, 0.95))
    h = capi.OfflineHandle(md, max_rows=8192, max_segments=256, batch_size=8, vad_dir=vd, vad_thres=thres)
    cut = h.vad_cut(pcm, 800, 15000)
    expect = []
    for b, e in cut:
        bs, es = min(int(b) * 16, len(pcm)), min(int(e) * 16, len(pcm))
        t = h.infer_segments(pcm, [bs], [es])
        expect.append(t[0] if isinstance(t, tuple) else t)
    h.close()
    expect = [e for e in expect if e.replace(" ", "")]        # a call whose segment decodes to nothing reports nothing
    assert len(expect) >= 3
    tp = capi.TpassStream(md, vd, options={"vad-speech-noise-thres": thres, "max-rows": 8192, "max-segments": 256})
    for chunk in (9600, 16000, 4000):                 # 600 ms (the server's chunking), 1 s, 250 ms
        conn = tp.connect()
        got = []
        for s in range(0, len(pcm), chunk):
            last = s + chunk >= len(pcm)
            r = tp.infer(conn, pcm[s:s + chunk], finished=last, mode=2, vad_tail_sil=800, vad_max_len=15000)
            assert r["msg"] == ""
            if r["tpass_msg"].replace(" ", ""):
                got.append(r["tpass_msg"])
        assert [g.replace(" ", "") for g in got] == [e.replace(" ", "") for e in expect], chunk
        # the connection is reusable after input_finished (Audio::ResetIndex): the same stream again gives the same texts
        r = tp.infer(conn, pcm[:200000], finished=True, mode=0)
        assert r["tpass_msg"] != "" or len(expect) == 0
    tp.close()


def test_whole_recording_vad_equals_reference_cutsplit(capi, synth, tmp_path):
    """The one structural deviation of the VAD path -- scores for the WHOLE recording in one pass + SegmentVad, instead of the
    reference's 1 s pieces through FsmnVadOnline (online fbank / LFR caches, FSMN caches through the session, incremental E2E
    scorer) -- checked against the reference's own compiled fsmn-vad.cpp / fsmn-vad-online.cpp / e2e-vad.h driven exactly like
    Audio::CutSplit (audio.cpp:1172-1226), with the oracle network behind the session: same frame count, same features, same
    segments."""
    import torch
    from oracle import am_ref as A
    if not A.available():
        pytest.skip("oracle/_ref/libfunasr_am_ref.so not built (needs /root/reference)")
    d = str(tmp_path)
    W, means, vars_ = synth.write_synthetic_vad_dir(d, seed=0)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    log = dict(feats=[])

    def net(ins):
        feats = ins[0][0]
        caches = [torch.from_numpy(c[0, :, :, 0].copy()) for c in ins[1:5]]
        assert all(c.shape == (1, 128, 19, 1) for c in ins[1:5])
        scores, nc = V.forward(feats, Wt, caches)
        log["feats"].append(feats.copy())
        return [scores.numpy()[None].astype(np.float32)] + [c.numpy()[None, :, :, None].astype(np.float32) for c in nc]

    rng = np.random.default_rng(3)
    n_seg_total = 0
    for case in range(6):
        parts = []
        for i in range(int(rng.integers(1, 6))):
            parts += [synth.make_audio(int(rng.integers(4000, 120000)), 300 + 10 * case + i), np.zeros(int(rng.integers(3000, 40000)), np.int16)]
        pcm = np.concatenate(parts)[: None if case % 2 else -int(rng.integers(1, 159))]
        if case == 5:
            pcm = pcm[:9000]                                    # shorter than one 1 s piece
        x = pcm.astype(np.float32) / np.float32(32768)
        feats = V.lfr_cmvn(F.fbank(x), means, vars_)
        scores, _ = V.forward(feats, Wt)
        p0 = scores.numpy()[:, 0]
        thres = float(np.clip(1.0 - 2.0 * np.quantile(p0, [0.5, 0.3, 0.7][case % 3]), 0.05, 0.95))
        ref = A.RefVad(d, net, speech_noise_thres=thres, tag="c%d" % case)
        for tail, mx in ((800, 15000), (250, 3000)):
            log["feats"] = []
            segs = ref.cut_split(x, tail, mx)
            seen = np.concatenate(log["feats"])
            assert seen.shape == feats.shape and np.abs(seen - feats).max() <= 1e-4       # frames and features of the chunked front end
            assert np.array_equal(segs, capi.host_vad_segments(p0, tail, mx, thres)), (case, tail, mx)
            n_seg_total += len(segs)
        ref.close()
    assert n_seg_total >= 10
