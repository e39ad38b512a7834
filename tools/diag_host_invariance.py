#!/usr/bin/env python
"""Host-library invariance on ONE GPU: Model::Forward on all segments of the 1 h stream vs on subsets (what a GPU's share of a
multi-GPU call is) -- the strings of the common segments must be identical."""
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
    tmp = tempfile.mkdtemp(prefix="b200pf_hinv_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    st = bench.make_stream()
    audio = [synth.make_audio(int(e - b), 900 + k).astype(np.float32) / np.float32(32768) for k, (b, e) in enumerate(st)]
    n = len(audio)
    h = capi.OfflineHandle(tmp, device=0, max_rows=65536, max_segments=4096, batch_size=4096)
    full = h.model_forward(audio)
    def sub(name, idx):
        r = h.model_forward([audio[i] for i in idx])
        bad = [i for k, i in enumerate(idx) if r[k] != full[i]]
        print("%-34s differing: %d %s" % (name, len(bad), [(i, len(full[i]), len(r[idx.index(i)])) for i in bad[:6]]), flush=True)
        for i in bad[:3]:
            mine = r[idx.index(i)]
            twins = [j for j in range(n) if full[j] == mine]
            same_len = [j for j in range(n) if len(audio[j]) == len(audio[i]) and j != i]
            a, b = full[i], mine
            k = next((t for t in range(min(len(a), len(b))) if a[t] != b[t]), min(len(a), len(b)))
            print("     seg %d: %d samples (T=%d); subset string equals full string of %s; segments of the same length %s; first difference at char %d of %d"
                  % (i, len(audio[i]), capi.lib().b200pf_num_lfr_frames(len(audio[i])), twins, same_len[:5], k, len(a)))
        return bad
    sub("all again", list(range(n)))
    sub("even indices", list(range(0, n, 2)))
    sub("odd indices", list(range(1, n, 2)))
    sub("first 100", list(range(100)))
    sub("first 40", list(range(40)))
    sub("100..306", list(range(100, n)))
    order = sorted(range(n), key=lambda i: -len(audio[i]))
    sub("LPT-like share (every 2nd longest)", sorted(order[0::2]))
    bad = sub("one at a time (first 60)", [0])
    cnt = 0
    for i in range(60):
        r = h.model_forward([audio[i]])
        if r[0] != full[i]:
            cnt += 1
            print("   alone differs:", i, len(audio[i]), len(full[i]), len(r[0]))
    print("alone differing among first 60:", cnt)
    h.close()


if __name__ == "__main__":
    main()
