"""Host logic of the cross-connection micro-batcher (csrc/host/micro_batcher.{h,cpp}) against a
host-only mock model: every caller gets ITS result, batches really form, the deadline and the size caps hold, segments
with different hotword matrices never share a batch, and batches are length-sorted like Audio::CutSplit's output."""
import threading
import time

import numpy as np


def _call(mb, n, first, out, i, hw=None):
    x = np.zeros(n, np.float32)
    if n:
        x[0] = first
    out[i] = mb.forward(x, hw)


def _parse(s):
    return dict(kv.split("=") for kv in s.split(";"))


def test_single_request_waits_for_the_deadline_only(capi):
    mb = capi.MicroBatcher(None, max_wait_us=30000, max_batch=64, max_rows=100000)
    t0 = time.time()
    r = mb.forward(np.full(16000, 7.0, np.float32))
    dt = time.time() - t0
    assert _parse(r) == {"n": "16000", "b": "1", "x": "7", "hw": "1"}
    assert 0.025 <= dt < 0.5
    st = mb.stats()
    assert st["segments"] == 1 and st["batches"] == 1 and st["closed_by_deadline"] == 1
    mb.close()


def test_concurrent_callers_are_batched_and_get_their_own_results(capi):
    mb = capi.MicroBatcher(None, max_wait_us=50000, max_batch=256, max_rows=10 ** 9, mock_latency_us=2000)
    n_thr = 48
    out = [None] * n_thr
    lens = [1600 * (1 + (7 * i) % 23) for i in range(n_thr)]
    th = [threading.Thread(target=_call, args=(mb, lens[i], i + 1, out, i)) for i in range(n_thr)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for i in range(n_thr):
        p = _parse(out[i])
        assert int(p["n"]) == lens[i] and int(p["x"]) == i + 1           # the caller's own segment came back
        assert int(p["b"]) > 1                                            # and it travelled in a real batch
    st = mb.stats()
    assert st["segments"] == n_thr and st["batches"] < n_thr / 4
    assert st["max_wait_us"] < 50000 + 100000
    assert "unsorted" not in "".join(out)                                 # ascending length inside each batch
    mb.close()


def test_size_caps_close_a_batch_early(capi):
    mb = capi.MicroBatcher(None, max_wait_us=2_000_000, max_batch=8, max_rows=10 ** 9)
    out = [None] * 16
    th = [threading.Thread(target=_call, args=(mb, 3200, i, out, i)) for i in range(16)]
    t0 = time.time()
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert time.time() - t0 < 1.5                                          # nobody waited for the 2 s deadline
    assert all(int(_parse(o)["b"]) == 8 for o in out)
    st = mb.stats()
    assert st["batches"] == 2 and st["closed_by_size"] == 2 and st["max_batch_seen"] == 8
    mb.close()
    # rows cap: 1 s = 17 LFR frames + 1 gap row = 18 rows; cap 40 rows -> at most 2 segments per batch
    mb = capi.MicroBatcher(None, max_wait_us=300000, max_batch=64, max_rows=40)
    out = [None] * 6
    th = [threading.Thread(target=_call, args=(mb, 16000, i, out, i)) for i in range(6)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert all(int(_parse(o)["b"]) <= 2 for o in out)
    mb.close()


def test_segments_with_different_hotwords_do_not_share_a_batch(capi):
    mb = capi.MicroBatcher(None, max_wait_us=100000, max_batch=64, max_rows=10 ** 9)
    hw_a = np.ones((3, 512), np.float32)
    hw_b = np.ones((5, 512), np.float32)
    hw_a2 = np.ones((3, 512), np.float32)                                   # equal content, other buffer: same batch as hw_a
    out = [None] * 9
    hws = [hw_a, hw_b, hw_a2] * 3
    th = [threading.Thread(target=_call, args=(mb, 1600 * (i + 1), i, out, i, hws[i])) for i in range(9)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for i, o in enumerate(out):
        p = _parse(o)
        assert int(p["hw"]) == hws[i].shape[0]
        assert int(p["b"]) == (3 if hws[i] is hw_b else 6)
    assert mb.stats()["batches"] == 2
    mb.close()


def test_destruction_drains_the_queue(capi):
    mb = capi.MicroBatcher(None, max_wait_us=5_000_000, max_batch=64, max_rows=10 ** 9)
    out = [None] * 4
    th = [threading.Thread(target=_call, args=(mb, 1600, i, out, i)) for i in range(4)]
    for t in th:
        t.start()
    time.sleep(0.2)
    mb.close()                                                              # callers are released with their results, not dropped
    for t in th:
        t.join(timeout=5)
    assert all(o is not None and o.startswith("n=1600") for o in out)


def test_multi_gpu_partition_is_balanced_and_complete(capi, synth):
    """MultiGpuParaformer::PartitionSegments: every segment gets exactly one queue; LPT keeps the per-queue FLOP estimate
    within a few percent; too-short segments are spread as well."""
    sch = __import__("importlib").import_module("asr-2pass_b200.scheduler")
    lens = synth.segment_lengths(1024).astype(np.int32)
    for n_dev in (2, 4, 8):
        a = capi.host_partition(lens, n_dev)
        assert a.min() == 0 and a.max() == n_dev - 1
        load = np.zeros(n_dev)
        for n, d in zip(lens, a):
            load[d] += sch.segment_cost(sch.num_lfr_frames(int(n)))
        assert load.max() / load.min() < 1.02
    assert list(capi.host_partition(np.array([16000, 100, 100, 100], np.int32), 2)) in ([0, 1, 1, 1], [0, 1, 1, 0], [0, 1, 0, 1])
    assert list(capi.host_partition(np.array([5, 6, 7], np.int32), 1)) == [0, 0, 0]
