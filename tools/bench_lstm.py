"""LSTM recurrence micro-benchmark:  python tools/bench_lstm.py"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
capi = importlib.import_module("asr-2pass_b200.capi")
cases = [(32, 1000, 1)] if os.environ.get('B200PF_LSTM_DBG') else [(32, 1000, 1), (32, 1000, 2), (128, 1000, 1), (256, 1000, 1), (512, 500, 2)]
for n_seq, length, n_dir in cases:
    ms, mc = capi.op_lstm_bench(n_seq, length, n_dir)
    tasks = ((n_seq + 31) // 32) * n_dir
    print("n_seq=%4d len=%5d dirs=%d tasks=%3d  %8.3f ms  -> %.2f us/step/wave-of-%d  (max active clusters %d)"
          % (n_seq, length, n_dir, tasks, ms, ms * 1e3 / length / max(1, -(-tasks // max(mc, 1))), mc, mc), flush=True)
