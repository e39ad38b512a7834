"""One forward of one full-size batch (Paraformer-large, ~24k packed rows) for ncu:  python tools/prof_run.py [rows]"""
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
    capi = importlib.import_module("asr-2pass_b200.capi")
    synth = importlib.import_module("asr-2pass_b200.synth")
    tmp = tempfile.mkdtemp(prefix="b200pf_prof_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    eng = capi.Engine(tmp, max_rows=rows, max_segments=4096)
    lens = synth.segment_lengths(1024)
    lens = np.sort(lens)[300:]  # mid-length segments, like a middle batch of the workload
    take, r = [], 0
    for n in lens:
        T = capi.lib().b200pf_num_lfr_frames(int(n)) + 1
        if r + T > rows:
            break
        take.append(int(n))
        r += T
    offs = np.concatenate([[0], np.cumsum(take)]).astype(np.int64)
    pcm = synth.make_audio(int(offs[-1]), 99)
    b = capi.Batch(eng, int(offs[-1]) + 64)
    res = b.forward_s16(pcm, offs)
    print("rows", r, "segments", len(take), "tokens", res["n_tokens"], "launches", b.launches)


if __name__ == "__main__":
    main()
