#!/usr/bin/env python
"""2-pass offline leg under many connections: C connection threads each decode their own closed VAD segments with
batch-1 Forward calls, as FunTpassInferBuffer does (funasrruntime.cpp:570-586) — once directly on the model, once
through the MicroBatcher.  Prints a JSON with both RTFx figures and the batcher's statistics (kept under profiles/).

    python tools/bench_microbatch.py [--connections 64] [--segments-per-connection 6]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--connections", type=int, default=64)
    ap.add_argument("--segments-per-connection", type=int, default=6)
    ap.add_argument("--max-wait-us", type=int, default=20000)
    args = ap.parse_args()
    synth = importlib.import_module("asr-2pass_b200.synth")
    capi = importlib.import_module("asr-2pass_b200.capi")
    tmp = tempfile.mkdtemp(prefix="b200pf_mb_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    h = capi.OfflineHandle(tmp, max_rows=32768, max_segments=1024, batch_size=256)
    C, S = args.connections, args.segments_per_connection
    lens = synth.segment_lengths(C * S, seed=4242)
    segs = [synth.make_audio(int(n), 10 + i).astype(np.float32) / np.float32(32768) for i, n in enumerate(lens)]
    audio_s = float(lens.sum()) / 16000.0
    out = {}

    def run(fn, name):
        res = [None] * (C * S)

        def conn(c):
            for k in range(S):
                res[c * S + k] = fn(segs[c * S + k])

        fn(segs[0])                                   # warm-up
        th = [threading.Thread(target=conn, args=(c,)) for c in range(C)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        out[name] = dict(wall_s=dt, rtfx=audio_s / dt)
        return res

    direct = run(lambda s: h.model_forward([s])[0], "direct_batch1")
    mb = capi.MicroBatcher(h, max_wait_us=args.max_wait_us, max_batch=256, max_rows=32768)
    batched = run(lambda s: mb.forward(s), "microbatched")
    out["microbatched"]["stats"] = mb.stats()
    # Vocab::Vector2StringV2 carries "previous call ended on an English word" across calls (vocab.cpp:177), so the call
    # ORDER decides a leading space; compare the texts without spaces
    strip = lambda xs: [x.replace(" ", "") for x in xs]
    out["identical_results"] = strip(direct) == strip(batched)
    out["config"] = "%d connections x %d segments U[2,20] s, Paraformer-large random-init, 1 B200, max_wait %d us" % (C, S, args.max_wait_us)
    out["audio_s"] = audio_s
    mb.close()
    h.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
