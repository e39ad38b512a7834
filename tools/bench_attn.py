"""Attention micro-benchmark on the configs[1] length distribution (and config 5's T = 1000):
python tools/bench_attn.py [n_segments] [iters]  -> one JSON line per case (ms per launch, TFLOP/s algorithmic)."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
capi = importlib.import_module("asr-2pass_b200.capi")
synth = importlib.import_module("asr-2pass_b200.synth")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    lens = synth.segment_lengths(n)
    T = np.sort(np.array([capi.lib().b200pf_num_lfr_frames(int(x)) for x in lens], np.int32))
    cases = [("configs1_self", T, False), ("configs1_cross", T, True), ("config5_self_T1000x32", np.full(32, 1000, np.int32), False),
             ("config5_cross_T1000x32", np.full(32, 1000, np.int32), True)]
    for name, t, cross in cases:
        ms, fl = capi.op_attention_bench(t, cross=cross, iters=iters)
        print(json.dumps({"case": name, "segments": int(len(t)), "rows": int(t.sum() + len(t)), "ms": round(ms, 4),
                          "tflops": round(fl / ms / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
