#!/usr/bin/env python
"""Timing of BASELINE.json configs[2] (contextual Paraformer + 100 hotwords + timestamp output) on one B200:
the first 256 segments of the config-2 workload, random-init full-size weights, device-resident PCM.  Not a
bench.py line (bench.py measures configs[1]); the JSON it prints is kept under profiles/.

    python tools/bench_config3.py [--segments 256] [--steps 3]
"""
import argparse
import importlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--segments", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--max-rows", type=int, default=65536)
    args = ap.parse_args()
    import torch
    synth = importlib.import_module("asr-2pass_b200.synth")
    capi = importlib.import_module("asr-2pass_b200.capi")
    mf = importlib.import_module("asr-2pass_b200.modelfile")
    bench = importlib.import_module("bench")
    pcm, offs = synth.make_segments(1024)
    lens = synth.segment_lengths(1024)[:args.segments]
    out = {}
    for name, extra in (("plain", {}), ("config3", dict(timestamp=1, contextual=1))):
        cfg, W = synth.make_weights(extra or None)
        means, vars_ = synth.make_cmvn()
        tmp = tempfile.mkdtemp(prefix="b200pf_c3_")
        mf.write_model_dir(tmp, cfg, W, means, vars_, synth.make_tokens(int(cfg["vocab"])))
        del W
        eng = capi.Engine(tmp, max_rows=args.max_rows, max_segments=4096)
        hw = None
        if extra:
            rng = np.random.default_rng(0)
            ids = np.zeros((101, 10), np.int32)
            ln = np.zeros(101, np.int32)
            for j in range(100):
                L = int(rng.integers(2, 7))
                ids[j, :L] = rng.integers(3, 8403, L)
                ln[j] = L
            ids[100, 0], ln[100] = 1, 1
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.hotword_embed(ids, ln)
            ev0.record()
            hw = eng.hotword_embed(ids, ln)
            ev1.record()
            ev1.synchronize()
            out["hotword_compile_ms"] = ev0.elapsed_time(ev1)
        groups = bench.make_batches(lens, args.max_rows, 4096, capi)
        batches = []
        for g in groups:
            buf = np.concatenate([pcm[offs[i]:offs[i + 1]] for i in g])
            ho = np.concatenate([[0], np.cumsum([lens[i] for i in g])]).astype(np.int64)
            b = capi.Batch(eng, len(buf) + 64)
            if hw is not None:
                b.set_hotwords(hw)
            b.stage_s16(buf, ho)
            batches.append(b)
        torch.cuda.synchronize()
        stream = torch.cuda.ExternalStream(eng.stream)
        for _ in range(2):
            for b in batches:
                b.run()
        res = [b.collect() for b in batches]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record()
            for _ in range(args.steps):
                for b in batches:
                    b.run()
            ev1.record()
        ev1.synchronize()
        ms = ev0.elapsed_time(ev1) / args.steps
        eng.set_option("profile", 1)
        eng.profile_read(reset=True)
        for b in batches:
            b.run()
        prof = eng.profile_read(reset=True)
        eng.set_option("profile", 0)
        audio_s = float(lens.sum()) / 16000.0
        out[name] = dict(ms_per_step=ms, rtfx=audio_s / (ms / 1e3), audio_s=audio_s, tokens=int(sum(r["n_tokens"] for r in res)),
                         launches=int(sum(b.launches for b in batches)),
                         kernels_ms={k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]})
        for b in batches:
            b.close()
        eng.close()
    out["config"] = "configs[2]: first %d segments of the config-2 workload, 100 hotwords + blank, timestamps on" % args.segments
    print(json.dumps(out))


if __name__ == "__main__":
    main()
