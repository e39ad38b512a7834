"""Ablation 64 of the attention kernel: clock64() stamps of CTA 0's first key blocks -> hop latencies (cycles).
B200PF_ATTN_DBG=64 [B200PF_ATTN_CTAS=1] python tools/trace_attn.py"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
capi = importlib.import_module("asr-2pass_b200.capi")
T = np.full(32, 1000, np.int32)
capi.op_attention_bench(T, iters=1)
buf = (C.c_longlong * 512)()
L = capi.lib()
L.b200pf_attn_trace_read.argtypes = [C.POINTER(C.c_longlong)]
assert L.b200pf_attn_trace_read(buf) == 0
t = np.array(buf[:], np.int64).reshape(64, 8)
t0 = t[0, 0]
print("block  S_commit softmax_sees_S softmax_arrives mma_sees_P | issue_s_enter K_there V_there PV_committed  (cycles since first commit)")
for g in range(2, 20):
    print(g, *(int(x - t0) for x in t[g]))
d_commit_to_seen = (t[4:40, 1] - t[4:40, 0])
d_softmax = (t[4:40, 2] - t[4:40, 1])
d_arrive_to_mma = (t[4:40, 3] - t[4:40, 2])
period = np.diff(t[4:40, 3])
print("median: S commit -> softmax sees S", int(np.median(d_commit_to_seen)), "| softmax work", int(np.median(d_softmax)),
      "| arrive -> MMA thread sees P", int(np.median(d_arrive_to_mma)), "| block period", int(np.median(period)))
