#!/usr/bin/env python
"""One invocation each of the two auxiliary engines, for ncu launch lists (profiles/r01_aux_launches.md):
    vad   one 10-minute recording through b200pf_vad_scores_s16            (python tools/prof_aux.py vad)
    punc  one lock-step round: 256 sequences of 40 tokens, full-size model   (python tools/prof_aux.py punc)
Prints the wall time of the second (warm) call."""
import importlib
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "vad"
    synth = importlib.import_module("asr-2pass_b200.synth")
    capi = importlib.import_module("asr-2pass_b200.capi")
    d = tempfile.mkdtemp(prefix="b200pf_aux_")
    if what == "vad":
        synth.write_synthetic_vad_dir(d, seed=0)
        eng = capi.VadEngine(d, max_frames=70000)
        pcm = np.tile(synth.make_audio(16000 * 60, 9), 10)
        offs = np.array([0, len(pcm)], np.int64)
        eng.scores(pcm, offs)
        t0 = time.perf_counter()
        p0, _, _, _ = eng.scores(pcm, offs)
        print("vad: %d frames in %.3f ms" % (len(p0), (time.perf_counter() - t0) * 1e3))
    else:
        cfg, W, toks = synth.write_synthetic_punc_dir(d, None, seed=0)
        eng = capi.PuncEngine(d, max_tokens=16384)
        rng = np.random.default_rng(0)
        n_seq, T = 256, 40
        ids = rng.integers(0, cfg["vocab"], n_seq * T).astype(np.int32)
        offs = (np.arange(n_seq + 1) * T).astype(np.int32)
        eng.infer(ids, offs)
        t0 = time.perf_counter()
        eng.infer(ids, offs)
        print("punc: %d tokens in %.3f ms, %d launches per call" % (len(ids), (time.perf_counter() - t0) * 1e3, eng.launches // 2))


if __name__ == "__main__":
    main()
