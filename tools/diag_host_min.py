#!/usr/bin/env python
"""Minimal reproduction hunt: segment 28 of the 1 h stream with different neighbours, host library vs engine."""
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
    tmp = tempfile.mkdtemp(prefix="b200pf_hmin_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    st = bench.make_stream()
    pcm16 = [synth.make_audio(int(e - b), 900 + k) for k, (b, e) in enumerate(st)]
    audio = [s.astype(np.float32) / np.float32(32768) for s in pcm16]
    h = capi.OfflineHandle(tmp, device=0, max_rows=65536, max_segments=4096, batch_size=4096)
    det = capi.Detokenizer(tmp) if hasattr(capi, "Detokenizer") else None
    target = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    alone = h.model_forward([audio[target]])[0]
    eng = capi.Engine(tmp, max_rows=65536, max_segments=4096)
    b = capi.Batch(eng, 40_000_000)
    r = b.forward_f32([audio[target]])
    ids_alone = r["token_ids"][:r["token_offsets"][1]].copy()
    combos = [[target - 2, target, target + 2], [target, target + 2], [target - 2, target], [target - 1, target, target + 1],
              list(range(0, 2 * 60, 2)), list(range(target - 20, target + 21, 2)), list(range(target - 6, target + 7, 2))]
    for c in combos:
        hs = h.model_forward([audio[i] for i in c])
        k = c.index(target)
        r = b.forward_f32([audio[i] for i in c])
        ids = r["token_ids"][r["token_offsets"][k]:r["token_offsets"][k + 1]]
        r2 = b.forward_s16(np.concatenate([pcm16[i] for i in c]), np.concatenate([[0], np.cumsum([len(pcm16[i]) for i in c])]).astype(np.int64))
        ids2 = r2["token_ids"][r2["token_offsets"][k]:r2["token_offsets"][k + 1]]
        print("neighbours %-40s host==alone %s | engine f32 ids==alone %s | engine s16 ids==alone %s" %
              (str(c[:7]) + ("..." if len(c) > 7 else ""), hs[k] == alone, np.array_equal(ids, ids_alone), np.array_equal(ids2, ids_alone)), flush=True)
    h.close()


if __name__ == "__main__":
    main()
