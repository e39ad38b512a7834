"""Reference-facing host API on the GPU (-m gpu): the funasrruntime.h call sequences of configs[0]
(FunOfflineInit -> FunOfflineInfer*/FunASRInfer -> FunASRGetResult) through libfunasr_b200.so, checked against the C ABI's
token ids run through the Python restatement of Vocab::Vector2StringV2 / the stitching rules."""
import os
import wave

import numpy as np
import pytest

from oracle import postproc_ref as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model(capi, synth, gpu, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("host"))
    cfg, W, means, vars_, toks = synth.write_synthetic_model_dir(d, dict(n_enc=2, n_dec=2), seed=0, jitter_ln=True)
    eng = capi.Engine(d, max_rows=4096, max_segments=64)
    return dict(dir=d, toks=toks, eng=eng)


def _ids(capi, model, pcm16):
    b = capi.Batch(model["eng"], len(pcm16) + 16)
    r = b.forward_s16(pcm16, np.array([0, len(pcm16)], np.int64))
    return list(r["token_ids"])


def _strip(s):
    return s.replace(" ", "")   # a leading space depends on Vocab's cross-call state (vocab.cpp:177), i.e. on call order


def test_config1_call_sequence_10s_wav(capi, synth, model, tmp_path):
    """configs[0]: one 10 s 16 kHz wav through FunOfflineInit(...,1,false,1) -> FunOfflineInfer -> FunASRGetResult."""
    pcm = synth.make_audio(160000, 42)
    h = capi.OfflineHandle(model["dir"], max_rows=4096, max_segments=64, batch_size=1)
    text, snippet = h.infer_buffer(pcm, vad_max_len=20000)
    assert abs(snippet - 10.0) < 1e-6
    expect = P.Vocab(model["toks"]).vector2string_v2(_ids(capi, model, pcm), "zh-cn")
    assert _strip(text) == _strip(expect) and len(text) > 0
    # the plain-model API on a wav file and on the raw buffer
    path = os.path.join(str(tmp_path), "a.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(pcm.astype("<i2").tobytes())
    t_file = capi.funasr_infer(model["dir"], wav_path=path, max_rows=4096)
    t_buf = capi.funasr_infer(model["dir"], pcm16=pcm, max_rows=4096)
    assert t_file == t_buf and _strip(t_file) == _strip(expect)
    h.close()


def test_hard_cut_and_segment_stitching(capi, synth, model):
    """No VAD in the stand-alone shim: audio longer than vad_max_len is hard-cut; texts of the pieces are concatenated in
    time order (funasrruntime.cpp:291-300); externally cut segments go through the FetchDynamic-style batching."""
    pcm = synth.make_audio(16000 * 25, 7)
    h = capi.OfflineHandle(model["dir"], max_rows=4096, max_segments=64, batch_size=8)
    text, snippet = h.infer_buffer(pcm, vad_max_len=10000)
    assert abs(snippet - 25.0) < 1e-6
    v = P.Vocab(model["toks"])
    pieces = [pcm[0:160000], pcm[160000:320000], pcm[320000:]]
    expect = "".join(v.vector2string_v2(_ids(capi, model, p), "zh-cn") for p in pieces)
    assert _strip(text) == _strip(expect)
    b, e = [320000, 0, 160000], [400000, 160000, 320000]                 # out of time order, different lengths
    seg_text = h.infer_segments(pcm, b, e)
    expect2 = "".join(v.vector2string_v2(_ids(capi, model, pcm[s:t]), "zh-cn") for s, t in zip(b, e))
    assert _strip(seg_text) == _strip(expect2)                           # results come back in the caller's segment order
    assert h.infer_buffer(np.zeros(0, np.int16))[0] == ""               # zero-length audio -> empty result object
    assert h.infer_buffer(np.zeros(300, np.int16))[0] == ""             # shorter than one fbank window -> ""
    h.close()


def test_init_failure_is_reported_not_fatal_through_the_hooks(capi, tmp_path):
    with pytest.raises(capi.B200PFError):
        capi.OfflineHandle(str(tmp_path))                                 # no model files
    with pytest.raises(capi.B200PFError):
        capi.funasr_infer(str(tmp_path), pcm16=np.zeros(16000, np.int16))


def test_concurrent_forward_calls_on_one_handle(capi, synth, model):
    """decoder-thread-num threads share one handle in the reference's servers (websocket-server.cpp:387-403): concurrent
    Forward calls must each get their own, correct result."""
    import threading
    h = capi.OfflineHandle(model["dir"], max_rows=4096, max_segments=64, batch_size=8)
    segs = [synth.make_audio(int(n), 8000 + i).astype(np.float32) / np.float32(32768)
            for i, n in enumerate([16000, 40000, 52800, 24000, 90000, 33000, 64000, 20000])]
    ref = [_strip(h.model_forward([s])[0]) for s in segs]
    out = [[None] * 6 for _ in segs]

    def worker(i):
        for k in range(6):
            out[i][k] = _strip(h.model_forward([segs[i], segs[(i + k) % len(segs)]])[0])

    th = [threading.Thread(target=worker, args=(i,)) for i in range(len(segs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for i in range(len(segs)):
        assert out[i] == [ref[i]] * 6
    h.close()


def test_reference_style_rtf_harness_runs(capi, synth, model, tmp_path):
    """tools/funasr_b200_offline_rtf.cpp (the reference's funasr-onnx-offline-rtf pattern against the shim): P threads on one
    handle, with and without the shim's micro-batcher; both must decode every wav."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "asr-2pass_b200", "lib", "funasr-b200-offline-rtf")
    assert os.path.exists(exe), "built by asr-2pass_b200/csrc/Makefile"
    scp = os.path.join(str(tmp_path), "wav.scp")
    with open(scp, "w") as f:
        for k, n in enumerate([16000, 52800, 33000, 24000, 80000, 20000]):
            path = os.path.join(str(tmp_path), "u%d.wav" % k)
            with wave.open(path, "wb") as w:
                w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
                w.writeframes(synth.make_audio(n, 70 + k).astype("<i2").tobytes())
            f.write("u%d %s\n" % (k, path))
    for extra in ([], ["--micro-batch-us", "10000"]):
        r = subprocess.run([exe, "--model-dir", model["dir"], "--wav-scp", scp, "--thread-num", "3", "--max-rows", "4096"] + extra,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-500:]
        j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        assert j["wavs"] == 6 and abs(j["audio_s"] - sum([16000, 52800, 33000, 24000, 80000, 20000]) / 16000.0) < 1e-2
        assert "speedup" in r.stdout and "total_rtf" in r.stdout


def test_reference_unmodified_rtf_harness_and_header_caller_run_on_the_gpu(capi, synth, model, tmp_path):
    """oracle/_ref/funasr-onnx-offline-rtf-ref is the reference's UNMODIFIED onnxruntime/bin/funasr-onnx-offline-rtf.cpp, and
    oracle/_ref/boundary_link_check a caller compiled against the reference's own funasrruntime.h -- both linked against
    libfunasr_b200.so in the build container (oracle/Makefile `boundary`, tests/test_boundary_cpu.py).  Here they RUN on the B200:
    the reference's own harness decodes a wav.scp with three threads on one handle and exits 0 (its LOG(INFO) lines go to the
    stand-in glog sink, so the exit status is the check); the header caller runs one FunOfflineInferBuffer request."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "funasr-onnx-offline-rtf-ref")
    chk = os.path.join(root, "oracle", "_ref", "boundary_link_check")
    if not (os.path.exists(exe) and os.path.exists(chk)):
        pytest.skip("oracle/_ref binaries are built in the container that holds /root/reference")
    scp = os.path.join(str(tmp_path), "wav.scp")
    with open(scp, "w") as f:
        for k, n in enumerate([16000, 52800, 33000, 24000]):
            path = os.path.join(str(tmp_path), "r%d.wav" % k)
            with wave.open(path, "wb") as w:
                w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
                w.writeframes(synth.make_audio(n, 90 + k).astype("<i2").tobytes())
            f.write("r%d %s\n" % (k, path))
    r = subprocess.run([exe, "--model-dir", model["dir"], "--wav-path", scp, "--thread-num", "3", "--quantize", "false"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout + r.stderr)[-800:]
    r = subprocess.run([chk, model["dir"]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "boundary_link_check ok" in r.stdout and "text_bytes=" in r.stdout, (r.stdout + r.stderr)[-800:]


def test_parallel_text_assembly_equals_the_serial_detokeniser(capi, synth, model):
    """Calls of >= 64 segments build every segment's text for both incoming detokeniser states on host threads and then follow
    the state chain; smaller calls run the reference's serial loop.  Same segments, same order, same previous call -> the same
    strings, spaces included (the state is carried from one call to the next exactly like Vocab::last_is_complete_english_)."""
    lens = [int(16000 * (1.0 + 0.13 * (i % 11))) for i in range(96)]
    segs = [synth.make_audio(n, 4100 + i).astype(np.float32) / np.float32(32768) for i, n in enumerate(lens)]
    big = capi.OfflineHandle(model["dir"], max_rows=4096, max_segments=128, batch_size=128)
    small = capi.OfflineHandle(model["dir"], max_rows=4096, max_segments=128, batch_size=128)
    for rnd in range(2):                                   # the second round starts from the state the first one left behind
        a = big.model_forward(segs)                        # 96 segments: parallel assembly
        b = []
        for k in range(0, 96, 32):                         # 32 at a time: the serial loop, same order
            b += small.model_forward(segs[k:k + 32])
        assert a == b, [i for i in range(96) if a[i] != b[i]][:5]
    big.close()
    small.close()
