#!/usr/bin/env python
"""Batch-1 latency of the 2pass-offline call (funasrruntime.cpp:570-586: one closed VAD segment -> Forward with batch 1):
p50 / p95 in milliseconds for segments of 2-20 s, full-size model, through
  (a) the C ABI (b200pf_forward_s16, pinned int16), CUDA graphs on and off,
  (b) funasr::Model::Forward(float*, int) of the host library (pageable float).
    python tools/bench_latency.py [n_calls]
"""
import importlib
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pct(v, p):
    v = sorted(v)
    return v[min(len(v) - 1, int(round(p / 100.0 * (len(v) - 1))))]


def main():
    n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    synth = importlib.import_module("asr-2pass_b200.synth")
    capi = importlib.import_module("asr-2pass_b200.capi")
    tmp = tempfile.mkdtemp(prefix="b200pf_lat_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    rng = np.random.default_rng(1)
    ten = synth.make_audio(160000, 5)
    mixed = [synth.make_audio(int(n), 100 + i) for i, n in enumerate(rng.integers(2 * 16000, 20 * 16000, 24))]
    out = {}
    for graphs in (1, 0):
        eng = capi.Engine(tmp, max_rows=8192, max_segments=64)
        eng.set_option("graphs", graphs)
        b = capi.Batch(eng, 16000 * 70)
        for case, segs in (("10s", [ten]), ("2-20s", mixed)):
            for s in segs * 3:                      # warm-up: eager pass, capture pass, replay
                b.forward_s16(s, np.array([0, len(s)], np.int64))
            lat = []
            ids0 = None
            for k in range(n_calls):
                s = segs[k % len(segs)]
                t0 = time.perf_counter()
                r = b.forward_s16(s, np.array([0, len(s)], np.int64))
                lat.append((time.perf_counter() - t0) * 1e3)
                if case == "10s":
                    ids0 = r["token_ids"].copy() if ids0 is None else ids0
                    assert np.array_equal(ids0, r["token_ids"])
            audio = sum(len(segs[k % len(segs)]) for k in range(n_calls)) / 16000.0
            out["c_abi_graphs%d_%s" % (graphs, case)] = dict(p50_ms=round(pct(lat, 50), 3), p95_ms=round(pct(lat, 95), 3), rtfx=round(audio / (sum(lat) / 1e3)),
                                                             launches=int(b.launches))
        out["graph_stats_graphs%d" % graphs] = eng.graph_stats()
        if graphs == 0:      # where the GPU time of a batch-1 forward goes: per-kernel-class CUDA-event time of the 10 s segment
            eng.set_option("profile", 1)
            eng.profile_read(reset=True)
            for _ in range(20):
                b.forward_s16(ten, np.array([0, len(ten)], np.int64))
            prof = eng.profile_read(reset=True)
            eng.set_option("profile", 0)
            out["profile_10s_us_per_launch"] = {k: round(1e3 * v["ms"] / max(1, v["launches"]), 2) for k, v in prof.items() if v["launches"]}
            out["profile_10s_ms_per_forward"] = {k: round(v["ms"] / 20, 3) for k, v in prof.items() if v["launches"]}
        b.close()
        eng.close()
    h = capi.OfflineHandle(tmp, max_rows=8192, max_segments=64, batch_size=1)
    f10 = ten.astype(np.float32) / np.float32(32768)
    for _ in range(5):
        h.model_forward([f10])
    lat = []
    for _ in range(n_calls):
        t0 = time.perf_counter()
        h.model_forward([f10])
        lat.append((time.perf_counter() - t0) * 1e3)
    out["model_forward_float_10s"] = dict(p50_ms=round(pct(lat, 50), 3), p95_ms=round(pct(lat, 95), 3), rtfx=round(10.0 * n_calls / (sum(lat) / 1e3)))
    h.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
