"""Multi-GPU host path (-m gpu; skipped unless the box has >= 2 B200s): one FunOfflineInit handle over two GPUs gives
the same strings as a single-GPU handle, and both GPUs take a share of every call (segments are
independent units, no collective)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_gpu_handle_equals_single_gpu(capi, synth, gpu, tmp_path_factory):
    if capi.device_count() < 2:
        pytest.skip("needs two B200s (gpurun --gpus 2)")
    d = str(tmp_path_factory.mktemp("mg"))
    synth.write_synthetic_model_dir(d, dict(n_enc=2, n_dec=2), seed=0, jitter_ln=True)
    segs = [synth.make_audio(int(n), 6000 + i).astype(np.float32) / np.float32(32768)
            for i, n in enumerate([16000, 52800, 33000, 8000, 120000, 20000, 64000, 300, 48000, 25000, 90000, 16000, 70000])]
    one = capi.OfflineHandle(d, max_rows=4096, max_segments=64, batch_size=64)
    ref = [one.model_forward([s])[0] for s in segs]        # per segment, so that Vocab's cross-call state cannot differ
    one.close()
    two = capi.OfflineHandle(d, max_rows=4096, max_segments=64, batch_size=64, devices=[0, 1])
    strip = lambda xs: [x.replace(" ", "") for x in xs]     # a leading space depends on the call ORDER (vocab.cpp:177)
    out = two.model_forward(segs)
    assert strip(out) == strip(ref)
    per = two.segments_per_device()
    assert len(per) == 2 and sum(per) == len(segs) and min(per) >= 3
    # FunOfflineInferBuffer over the pool, and the micro-batcher in front of it
    pcm = synth.make_audio(16000 * 30, 5)
    text, _ = two.infer_buffer(pcm, vad_max_len=10000)      # 3 hard-cut segments over 2 GPUs
    single = capi.OfflineHandle(d, max_rows=4096, max_segments=64, batch_size=64)
    text1, _ = single.infer_buffer(pcm, vad_max_len=10000)
    assert text.replace(" ", "") == text1.replace(" ", "")
    single.close()
    import threading
    mb = capi.MicroBatcher(two, max_wait_us=30000, max_batch=64, max_rows=4096)
    got = [None] * len(segs)
    th = [threading.Thread(target=lambda i=i: got.__setitem__(i, mb.forward(segs[i]))) for i in range(len(segs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert strip(got) == strip(ref)
    mb.close()
    two.close()


def test_two_gpu_handle_returns_the_one_gpu_text_exactly(capi, synth, gpu, tmp_path_factory):
    """Spaces included: the pool assembles every call's text in the caller's order through ONE detokeniser state chain (the
    reference's Vocab carries a leading-space decision from call to call), so for the same sequence of calls a two-GPU handle and
    a one-GPU handle return identical strings -- sharding changes no character."""
    if capi.device_count() < 2:
        pytest.skip("needs two B200s (gpurun --gpus 2)")
    d = str(tmp_path_factory.mktemp("mg2"))
    synth.write_synthetic_model_dir(d, dict(n_enc=2, n_dec=2), seed=0, jitter_ln=True)
    lens = [int(16000 * (1.0 + 0.17 * (i % 13))) for i in range(80)]
    segs = [synth.make_audio(n, 9100 + i).astype(np.float32) / np.float32(32768) for i, n in enumerate(lens)]
    one = capi.OfflineHandle(d, max_rows=4096, max_segments=128, batch_size=128)
    two = capi.OfflineHandle(d, max_rows=4096, max_segments=128, batch_size=128, devices=[0, 1])
    for call in (segs, segs[10:30], segs[::3], segs):          # the same sequence of calls on both handles
        assert two.model_forward(call) == one.model_forward(call)
    one.close()
    two.close()
