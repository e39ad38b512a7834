#!/usr/bin/env python
"""Where the time of funasr::Model::Forward(float**, int*) goes on the configs[1] workload: the float staging alone
(b200pf_batch_stage_f32: exact float -> int16 into pinned memory with host threads + one H2D) and the whole call for several
sub-batch caps (B200PF_SUB_ROWS; unset = geometric growth from 8192 rows).   python tools/bench_model_forward.py"""
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(sub_rows, max_rows):
    synth = importlib.import_module("asr-2pass_b200.synth")
    capi = importlib.import_module("asr-2pass_b200.capi")
    tmp = tempfile.mkdtemp(prefix="b200pf_mf_")
    synth.write_synthetic_model_dir(tmp, None, seed=0)
    pcm, offs = synth.make_segments(1024)
    lens = synth.segment_lengths(1024)
    order = np.argsort(lens, kind="stable")
    fsegs = [pcm[offs[i]:offs[i + 1]].astype(np.float32) / np.float32(32768) for i in order]
    audio_s = float(lens.sum()) / 16000.0
    out = dict(sub_rows=sub_rows, max_rows=max_rows)
    if sub_rows == "stage":
        eng = capi.Engine(tmp, max_rows=196608, max_segments=4096)
        b = capi.Batch(eng, int(lens.sum()) + 64)
        for _ in range(2):
            b.stage_f32(fsegs)
        import ctypes
        t0 = time.perf_counter()
        for _ in range(3):
            b.stage_f32(fsegs)
        ctypes.CDLL(None)  # no-op
        eng_stream_sync = capi.lib().b200pf_batch_run  # keep the symbol alive
        b.run(); b.collect()
        out["stage_f32_ms"] = (time.perf_counter() - t0) / 3 * 1e3
    else:
        h = capi.OfflineHandle(tmp, max_rows=max_rows, max_segments=4096, batch_size=4096)
        h.model_forward(fsegs[:64])
        h.model_forward(fsegs)
        t0 = time.perf_counter()
        for _ in range(4):
            h.model_forward(fsegs)
        dt = (time.perf_counter() - t0) / 4
        out.update(ms=dt * 1e3, rtfx=audio_s / dt)
        h.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1], int(sys.argv[2]))
    else:
        for sub, mr in (("stage", 0), ("0", 65536), ("0", 196608), ("65536", 65536), ("98304", 196608), ("196608", 196608)):
            env = dict(os.environ)
            if sub not in ("stage", "0"):
                env["B200PF_SUB_ROWS"] = sub
            subprocess.run([sys.executable, os.path.abspath(__file__), sub, str(mr)], env=env)
