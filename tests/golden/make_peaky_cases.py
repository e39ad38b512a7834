"""Generator of tests/golden/peaky_cases.json (run once in the build container; CPU only):

    python tests/golden/make_peaky_cases.py

Searches seeded short segments on the small "peaky" synthetic model (synth.make_peaky: only 64 vocabulary entries reachable)
for cases in which EVERY row's top-1 margin, every CIF integrate value and the token count clear their decision thresholds by
far more than the CUDA path's rounding error (margins stored in the file).  On those cases token ids, fire frames and token
counts must equal the fp32 oracle's exactly -- no tie rule applies (tests/test_gpu_parity.py::test_peaky_cases_are_bit_exact).
A random-init network cannot be made tie-free in general (the top-2 gap of ~90 rows x 64 classes and ~100 integrate values per
segment fall near a threshold somewhere with high probability), hence the search; everything else is covered by the
margin-aware rules and mismatch-rate bounds of tests/test_gpu_parity_fullsize.py.
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
synth = importlib.import_module("asr-2pass_b200.synth")
from oracle import frontend as F  # noqa: E402
from oracle import paraformer_ref as R  # noqa: E402

MIN_GAP, MIN_FIRE, MIN_TOK, WANT, TRIALS = 0.05, 0.01, 0.02, 5, 1500


def main():
    cfgo = dict(n_enc=2, n_dec=2)
    cfg, W = synth.make_weights(cfgo, 0, True)
    synth.make_peaky(W)
    means, vars_ = synth.make_cmvn(560)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    pc = R.PfConfig.from_dict(cfg)
    rng = np.random.default_rng(123)
    found = []
    for trial in range(TRIALS):
        n = int(rng.integers(1 * 16000, 5 * 16000))
        seed = 9000 + trial
        pcm = synth.make_audio(n, seed)
        feats = F.lfr_cmvn(F.fbank(pcm.astype(np.float32) / np.float32(32768)), means, vars_)
        o = R.forward(feats, Wt, pc, want_taps=False)
        lg = o["logits"].numpy()
        if lg.shape[0] == 0:
            continue
        top2 = np.sort(lg, axis=1)[:, -2:]
        gap = float((top2[:, 1] - top2[:, 0]).min())
        fires = o["fires"].numpy()
        fmargin = float(np.abs(fires - 1.0).min())
        s = float(o["alphas"].double().sum())
        tmargin = float(min(s - np.floor(s), np.ceil(s) - s))
        if gap >= MIN_GAP and fmargin >= MIN_FIRE and tmargin >= MIN_TOK:
            found.append(dict(n_samples=n, audio_seed=seed, T=int(feats.shape[0]), L=len(o["ids"]), min_gap=gap, fire_margin=fmargin,
                              tok_margin=tmargin, ids=[int(i) for i in o["ids"]], fire_frames=[int(t) for t in np.where(fires >= 1)[0]]))
            print(trial, n, seed, found[-1]["T"], found[-1]["L"], gap, fmargin, tmargin, flush=True)
        if len(found) >= WANT:
            break
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "peaky_cases.json")
    json.dump(dict(model=dict(cfg=cfgo, seed=0, jitter_ln=True, transform="make_peaky"), cases=found), open(out, "w"))
    print(len(found), "cases ->", out)


if __name__ == "__main__":
    main()
