#!/usr/bin/env python
"""Batch invariance at FULL size: the same segments decoded (A) in one batch, (B) in batches of 16, (C) one at a time
(the CUDA-graph path), (D) reversed order in one batch, (E) run A again.  Prints, per comparison, how many segments differ in ids / fire frames
and the largest difference of the encoder output, alphas and logits taps.
    python tools/diag_invariance.py [n_segments]
"""
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
capi = importlib.import_module("asr-2pass_b200.capi")
synth = importlib.import_module("asr-2pass_b200.synth")


def run(eng, b, segs, order, group):
    out = {}
    for g0 in range(0, len(order), group):
        idx = order[g0:g0 + group]
        pcm = np.concatenate([segs[i] for i in idx])
        offs = np.concatenate([[0], np.cumsum([len(segs[i]) for i in idx])]).astype(np.int64)
        r = b.forward_s16(pcm, offs)
        for k, i in enumerate(idx):
            a, e = r["token_offsets"][k], r["token_offsets"][k + 1]
            out[i] = dict(ids=r["token_ids"][a:e].copy(), fires=r["fire_frames"][a:e].copy(), enc=b.tap("enc", k), alphas=b.tap("alphas", k),
                          logits=b.tap("logits", k))
    return out


def cmp(name, x, y):
    n_ids = n_fire = 0
    d = dict(enc=0.0, alphas=0.0, logits=0.0)
    for i in x:
        n_ids += int(len(x[i]["ids"]) != len(y[i]["ids"]) or not np.array_equal(x[i]["ids"], y[i]["ids"]))
        n_fire += int(len(x[i]["fires"]) != len(y[i]["fires"]) or not np.array_equal(x[i]["fires"], y[i]["fires"]))
        for k in d:
            if x[i][k].shape == y[i][k].shape and x[i][k].size:
                d[k] = max(d[k], float(np.abs(x[i][k] - y[i][k]).max()))
            elif x[i][k].shape != y[i][k].shape:
                d[k] = float("inf")
    print("%-28s segments with different ids %d, fires %d | max |d| enc %.3g alphas %.3g logits %.3g" % (name, n_ids, n_fire, d["enc"], d["alphas"], d["logits"]), flush=True)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    tmp = tempfile.mkdtemp(prefix="b200pf_inv_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    eng = capi.Engine(tmp, max_rows=16384, max_segments=256)
    eng.set_option("taps", 1)
    lens = synth.segment_lengths(1024)[:: max(1, 1024 // n)][:n]
    segs = [synth.make_audio(int(x), 500 + i) for i, x in enumerate(lens)]
    b = capi.Batch(eng, int(sum(lens)) + 64)
    order = list(np.argsort(lens, kind="stable"))
    A = run(eng, b, segs, order, len(order))
    E = run(eng, b, segs, order, len(order))
    cmp("A vs A again", A, E)
    B = run(eng, b, segs, order, 16)
    cmp("one batch vs batches of 16", A, B)
    D = run(eng, b, segs, order[::-1], len(order))
    cmp("one batch vs reversed order", A, D)
    Cc = run(eng, b, segs, order, 1)
    cmp("one batch vs one at a time", A, Cc)
    C2 = run(eng, b, segs, order, 1)
    cmp("one at a time, twice", Cc, C2)
    eng.set_option("graphs", 0)
    C3 = run(eng, b, segs, order, 1)
    cmp("one at a time, graphs off", Cc, C3)
    eng.set_option("overlap", 0)
    C4 = run(eng, b, segs, order, 1)
    cmp("one at a time, no forks", C3, C4)
    eng.set_option("ffn_ln_fold", 0)
    A5 = run(eng, b, segs, order, len(order))
    B5 = run(eng, b, segs, order, 16)
    cmp("no LN fold: 1 batch vs 16s", A5, B5)


if __name__ == "__main__":
    main()
