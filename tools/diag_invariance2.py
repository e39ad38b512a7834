#!/usr/bin/env python
"""Batch invariance on the 1 h stream's segments in ARRIVAL order (mixed lengths side by side): one batch vs halves vs alone."""
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
diag = importlib.import_module("tools.diag_invariance") if False else None
sys.path.insert(0, os.path.join(ROOT, "tools"))
import diag_invariance as D


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    tmp = tempfile.mkdtemp(prefix="b200pf_inv2_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    eng = capi.Engine(tmp, max_rows=65536, max_segments=1024)
    eng.set_option("taps", 1)
    st = bench.make_stream()[:n]
    segs = [synth.make_audio(int(e - b), 900 + k) for k, (b, e) in enumerate(st)]
    b = capi.Batch(eng, int(sum(len(s) for s in segs)) + 64)
    order = list(range(n))
    A = D.run(eng, b, segs, order, n)
    H = D.run(eng, b, segs, order, (n + 1) // 2)
    D.cmp("arrival order: 1 batch vs halves", A, H)
    Q = D.run(eng, b, segs, order, 16)
    D.cmp("1 batch vs batches of 16", A, Q)
    S = D.run(eng, b, segs, sorted(order, key=lambda i: len(segs[i])), n)
    D.cmp("arrival order vs sorted", A, S)
    for i in order:
        if len(A[i]["ids"]) != len(Q[i]["ids"]) or not np.array_equal(A[i]["ids"], Q[i]["ids"]):
            print("segment", i, "samples", len(segs[i]), "tokens", len(A[i]["ids"]), len(Q[i]["ids"]),
                  "enc diff", float(np.abs(A[i]["enc"] - Q[i]["enc"]).max()), "alpha diff", float(np.abs(A[i]["alphas"] - Q[i]["alphas"]).max()))
            d = np.abs(A[i]["enc"] - Q[i]["enc"]).max(axis=1)
            print("   first rows where enc differs:", np.nonzero(d)[0][:10], "of", len(d))
            break


if __name__ == "__main__":
    main()
