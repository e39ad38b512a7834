#!/usr/bin/env python
"""BASELINE.json configs[3]: the 2pass-offline correction path on a 1 h synthetic stream.  Speech bursts U[2,20] s
separated by silences U[0.3,1.0] s (rng 4242); the host FSMN-VAD is not part of this path, so the segment boundaries are
the generator's ground truth.  The segments go, in arrival order, through ONE handle over N GPUs
(FunOfflineInferSegmentsB200 -> MultiGpuParaformer: length sort, FetchDynamic-style batches, per-GPU queues) and, second
variant, as per-connection batch-1 calls through the MicroBatcher.  AM-only RTFx = stream seconds / wall seconds.

    python tools/bench_config4.py --gpus 2
"""
import argparse
import importlib
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_stream(seconds=3600.0, seed=4242):
    rng = np.random.default_rng(seed)
    t, segs = 0.0, []
    while True:
        t += float(rng.uniform(0.3, 1.0))
        d = round(float(rng.uniform(2.0, 20.0)) * 100.0) / 100.0
        if t + d > seconds:
            break
        segs.append((int(t * 16000), int((t + d) * 16000)))
        t += d
    return segs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    synth = importlib.import_module("asr-2pass_b200.synth")
    capi = importlib.import_module("asr-2pass_b200.capi")
    tmp = tempfile.mkdtemp(prefix="b200pf_c4_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    segs = make_stream()
    n_samples = 3600 * 16000
    pcm = np.zeros(n_samples, np.int16)
    for k, (b, e) in enumerate(segs):
        if k % 16 == 0:
            blk = synth.make_audio(21 * 16000 * 16, 77 + k)       # one generation pass per 16 segments
        o = (k % 16) * 21 * 16000
        pcm[b:e] = blk[o:o + (e - b)]
    speech_s = sum(e - b for b, e in segs) / 16000.0
    out = dict(config="configs[3]: 1 h stream, %d ground-truth VAD segments (%.0f s of speech), arrival order" % (len(segs), speech_s))
    for n in sorted({1, args.gpus}):
        h = capi.OfflineHandle(tmp, max_rows=65536, max_segments=4096, batch_size=256, devices=list(range(n)))
        b = [s[0] for s in segs]
        e = [s[1] for s in segs]
        h.infer_segments(pcm, b[:32], e[:32])
        h.infer_segments(pcm, b, e)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            text = h.infer_segments(pcm, b, e)
        dt = (time.perf_counter() - t0) / args.steps
        r = dict(offline_call=dict(wall_s=dt, rtfx_stream=3600.0 / dt, rtfx_speech=speech_s / dt, chars=len(text)))
        # the same segments as batch-1 calls of 64 "connections" through the micro-batcher
        mb = capi.MicroBatcher(h, max_wait_us=20000, max_batch=256, max_rows=32768)
        fs = [pcm[s:t].astype(np.float32) / np.float32(32768) for s, t in segs]
        C = 64

        def conn(c):
            for k in range(c, len(fs), C):
                mb.forward(fs[k])

        mb.forward(fs[0])
        th = [threading.Thread(target=conn, args=(c,)) for c in range(C)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        r["microbatched_64_connections"] = dict(wall_s=dt, rtfx_stream=3600.0 / dt, rtfx_speech=speech_s / dt, stats=mb.stats())
        mb.close()
        h.close()
        out["gpus_%d" % n] = r
    print(json.dumps(out))


if __name__ == "__main__":
    main()
