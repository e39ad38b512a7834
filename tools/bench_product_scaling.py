#!/usr/bin/env python
"""Strong scaling of the PRODUCT's one-handle multi-GPU path (FunOfflineInferSegmentsB200 -> MultiGpuParaformer) on one box:
BASELINE.json configs[3] (1 h stream) and configs[4] (256 x 60 s) for N = 1, 2, 4, 8 GPUs (as many as the box has), the same
legs bench.py prints as config4 / config5, measured back to back so that the ratios are not box-to-box clock noise.
    python tools/bench_product_scaling.py [steps]
"""
import importlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    bench = importlib.import_module("bench")
    capi = importlib.import_module("asr-2pass_b200.capi")
    synth = importlib.import_module("asr-2pass_b200.synth")
    tmp = tempfile.mkdtemp(prefix="b200pf_scale_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    n_dev = capi.device_count()
    base = {}
    for n in (1, 2, 4, 8):
        if n > n_dev:
            break
        r = bench.product_multigpu(capi, synth, tmp, n, steps)
        for k in ("config4", "config5"):
            base.setdefault(k, r[k]["value"])
            r[k]["speedup_vs_1gpu"] = round(r[k]["value"] / base[k], 3)
        print(json.dumps(dict(n_gpus=n, config4=r["config4"], config5=r["config5"])), flush=True)


if __name__ == "__main__":
    main()
