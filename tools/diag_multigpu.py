#!/usr/bin/env python
"""Where do the strings of a 2-GPU handle differ from a 1-GPU handle's?  Full-size model, the 1 h stream's 307 segments as float
segments through Model::Forward on: GPU 0 alone (twice), GPU 1 alone, both GPUs (twice)."""
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
capi = importlib.import_module("asr-2pass_b200.capi")
synth = importlib.import_module("asr-2pass_b200.synth")
bench = importlib.import_module("bench")


def main():
    tmp = tempfile.mkdtemp(prefix="b200pf_mg_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    segs = bench.make_stream()
    audio = [synth.make_audio(int(e - b), 900 + k).astype(np.float32) / np.float32(32768) for k, (b, e) in enumerate(segs)]
    def run(devs, **kw):
        h = capi.OfflineHandle(tmp, max_rows=65536, max_segments=4096, batch_size=4096, **kw) if devs is None else \
            capi.OfflineHandle(tmp, max_rows=65536, max_segments=4096, batch_size=4096, devices=devs)
        a = h.model_forward(audio)
        b = h.model_forward(audio)
        h.close()
        return a, b
    g0a, g0b = run(None, device=0)
    g1a, g1b = run(None, device=1)
    ma, mb = run([0, 1])
    def diff(n, x, y):
        d = [i for i in range(len(x)) if x[i] != y[i]]
        print("%-26s differing segments: %d %s" % (n, len(d), [(i, len(audio[i]), len(x[i]), len(y[i])) for i in d[:6]]), flush=True)
    diff("gpu0 run1 vs run2", g0a, g0b)
    diff("gpu1 run1 vs run2", g1a, g1b)
    diff("gpu0 vs gpu1", g0a, g1a)
    diff("2-gpu run1 vs run2", ma, mb)
    diff("2-gpu vs gpu0", ma, g0a)


if __name__ == "__main__":
    main()
