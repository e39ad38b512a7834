#!/usr/bin/env python
"""One FunOfflineInit handle over N GPUs of the box (funasr_b200::MultiGpuParaformer): the configs[1] workload through
funasr::Model::Forward(float**, int*, ...) in ONE call, host float buffers in, strings out.  Prints a JSON.

    python tools/bench_multigpu_host.py --gpus 2
"""
import argparse
import importlib
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--segments", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    synth = importlib.import_module("asr-2pass_b200.synth")
    capi = importlib.import_module("asr-2pass_b200.capi")
    tmp = tempfile.mkdtemp(prefix="b200pf_mg_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    pcm, offs = synth.make_segments(args.segments)
    segs = [pcm[offs[i]:offs[i + 1]].astype(np.float32) / np.float32(32768) for i in range(args.segments)]
    audio_s = float(offs[-1]) / 16000.0
    out = dict(config="configs[1] through Model::Forward on one handle, float host buffers, %d segments" % args.segments, audio_s=audio_s)
    for n in sorted({1, args.gpus}):
        h = capi.OfflineHandle(tmp, max_rows=65536, max_segments=4096, batch_size=4096, devices=list(range(n)))
        h.model_forward(segs[:64])
        h.model_forward(segs)                                  # warm-up at full size (workspace, batch buffers)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = h.model_forward(segs)
        dt = (time.perf_counter() - t0) / args.steps
        out["gpus_%d" % n] = dict(wall_s=dt, rtfx=audio_s / dt, segments_per_device=h.segments_per_device(), nonempty=sum(1 for r in res if r))
        h.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
