#!/usr/bin/env python
"""Engine level, one GPU: the even-indexed segments of the 1 h stream as ONE call through (a) stage_s16 (contiguous int16),
(b) stage_f32 (float -> exact int16 into pinned staging, aligned starts), each compared with the segments decoded alone."""
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


def ids_of(r, k):
    return r["token_ids"][r["token_offsets"][k]:r["token_offsets"][k + 1]].copy()


def main():
    tmp = tempfile.mkdtemp(prefix="b200pf_sp_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    eng = capi.Engine(tmp, max_rows=65536, max_segments=4096)
    eng.set_option("taps", 1)
    st = bench.make_stream()
    segs = [synth.make_audio(int(e - b), 900 + k) for k, (b, e) in enumerate(st)]
    idx = list(range(0, len(segs), 2))[:120]
    b = capi.Batch(eng, int(sum(len(segs[i]) for i in idx)) + 4096)
    alone = {}
    for i in idx:
        r = b.forward_s16(segs[i], np.array([0, len(segs[i])], np.int64))
        alone[i] = (ids_of(r, 0), b.tap("fbank", 0), b.tap("enc", 0))
    pcm = np.concatenate([segs[i] for i in idx])
    offs = np.concatenate([[0], np.cumsum([len(segs[i]) for i in idx])]).astype(np.int64)
    for name, fn in (("stage_s16 contiguous", lambda: b.forward_s16(pcm, offs)),
                     ("stage_f32 exact->int16", lambda: b.forward_f32([segs[i].astype(np.float32) / np.float32(32768) for i in idx]))):
        r = fn()
        bad = []
        for k, i in enumerate(idx):
            if not np.array_equal(ids_of(r, k), alone[i][0]):
                fb, enc = b.tap("fbank", k), b.tap("enc", k)
                dfb = np.abs(fb - alone[i][1]).max(axis=1)
                bad.append((i, len(segs[i]), float(dfb.max()), np.nonzero(dfb)[0][:4].tolist(), len(dfb), float(np.abs(enc - alone[i][2]).max())))
        print("%-26s differing: %d" % (name, len(bad)))
        for x in bad[:8]:
            print("   seg %d samples %d | fbank max diff %.3g at frames %s of %d | enc max diff %.3g" % x)


if __name__ == "__main__":
    main()
