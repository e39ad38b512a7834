#!/usr/bin/env python
"""The host library's call pattern at engine level (taps off): sub-batches through two alternating batch objects, batch i+1 staged
(copy stream) and ENQUEUED before batch i is collected -- against each sub-batch run and collected on its own, and against other
splits of the same segments.  Segments = the 1 h stream's, arrival order."""
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


def pack(segs, idx):
    pcm = np.concatenate([segs[i] for i in idx])
    offs = np.concatenate([[0], np.cumsum([len(segs[i]) for i in idx])]).astype(np.int64)
    return pcm, offs


def split_rows(segs, idx, caps):
    """sub-batches by a growing row cap like ParaformerB200::RunAll"""
    out, cur, rows, k = [], [], 0, 0
    for i in idx:
        r = capi.lib().b200pf_num_lfr_frames(len(segs[i])) + 1
        if cur and rows + r > caps[min(k, len(caps) - 1)]:
            out.append(cur); cur = []; rows = 0; k += 1
        cur.append(i); rows += r
    if cur:
        out.append(cur)
    return out


def serial(eng, b, segs, groups):
    res = {}
    for g in groups:
        r = b.forward_f32([segs[i].astype(np.float32) / np.float32(32768) for i in g]) if os.environ.get("STAGE") == "f32" else b.forward_s16(*pack(segs, g))
        for k, i in enumerate(g):
            res[i] = (r["token_ids"][r["token_offsets"][k]:r["token_offsets"][k + 1]].copy(), r["fire_frames"][r["token_offsets"][k]:r["token_offsets"][k + 1]].copy())
    return res


def pipelined(eng, slots, segs, groups, copy_stream):
    res = {}
    keep = []
    def launch(j):
        p = pack(segs, groups[j]); keep.append(p)
        if os.environ.get("STAGE") == "f32":
            fs = [segs[i].astype(np.float32) / np.float32(32768) for i in groups[j]]
            keep.append(fs)
            slots[j & 1].stage_f32(fs, stream=copy_stream)
        else:
            slots[j & 1].stage_s16(p[0], p[1], stream=copy_stream)
        slots[j & 1].run()
    launch(0)
    for j in range(len(groups)):
        if j + 1 < len(groups):
            launch(j + 1)
        r = slots[j & 1].collect()
        for k, i in enumerate(groups[j]):
            res[i] = (r["token_ids"][r["token_offsets"][k]:r["token_offsets"][k + 1]].copy(), r["fire_frames"][r["token_offsets"][k]:r["token_offsets"][k + 1]].copy())
    return res


def cmp(name, x, y):
    bad = [i for i in x if len(x[i][0]) != len(y[i][0]) or not np.array_equal(x[i][0], y[i][0]) or not np.array_equal(x[i][1], y[i][1])]
    print("%-44s differing segments: %d %s" % (name, len(bad), bad[:8]), flush=True)


def main():
    tmp = tempfile.mkdtemp(prefix="b200pf_inv3_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    eng = capi.Engine(tmp, max_rows=65536, max_segments=4096)
    st = bench.make_stream()
    segs = [synth.make_audio(int(e - b), 900 + k) for k, (b, e) in enumerate(st)]
    n = len(segs)
    cap = int(sum(len(s) for s in segs)) + 64
    slots = [capi.Batch(eng, cap), capi.Batch(eng, cap)]
    all_idx = list(range(n))
    g1 = split_rows(segs, all_idx, [8192, 32768, 65536])
    even, odd = all_idx[0::2], all_idx[1::2]
    g2 = split_rows(segs, even, [8192, 32768, 65536]) + split_rows(segs, odd, [8192, 32768, 65536])
    L = capi.lib()
    L.b200pf_engine_copy_stream.argtypes = [__import__("ctypes").c_void_p]
    L.b200pf_engine_copy_stream.restype = __import__("ctypes").c_void_p
    cs = L.b200pf_engine_copy_stream(eng.h)
    print("sub-batch sizes", [len(g) for g in g1], "|", [len(g) for g in g2])
    for ov in (2, 0):
        eng.set_option("overlap", ov)
        A = serial(eng, slots[0], segs, g1)
        B = serial(eng, slots[0], segs, g2)
        cmp("overlap %d: serial split-1 vs split-2" % ov, A, B)
        P = pipelined(eng, slots, segs, g1, cs)
        cmp("overlap %d: serial vs pipelined (split-1)" % ov, A, P)
        P2 = pipelined(eng, slots, segs, g2, cs)
        cmp("overlap %d: serial vs pipelined (split-2)" % ov, B, P2)
        O = serial(eng, slots[0], segs, [[i] for i in all_idx[:64]])
        cmp("overlap %d: split-1 vs one at a time (64)" % ov, {i: A[i] for i in all_idx[:64]}, O)


if __name__ == "__main__":
    main()
