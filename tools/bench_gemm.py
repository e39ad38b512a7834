"""GEMM micro-benchmark at the encoder's shapes:  python tools/bench_gemm.py [rows]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
capi = importlib.import_module("asr-2pass_b200.capi")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
tot_ms = 0.0
for name, N, K, mode in [("qkv", 1536, 512, 0), ("out+mem+res", 512, 512, 3), ("ffn1 relu", 2048, 512, 1), ("ffn2 res", 512, 2048, 2),
                         ("dec kv", 1024, 512, 0), ("vocab argmax (M/2)", 8404, 512, 4)]:
    m = M // 2 if "vocab" in name else M
    ms = capi.op_gemm_bench(m, N, K, mode, 20)
    if name in ("qkv", "out+mem+res", "ffn1 relu", "ffn2 res"):
        tot_ms += ms
    print("%-20s M=%6d N=%5d K=%5d  %8.1f us  %7.1f TFLOP/s" % (name, m, N, K, ms * 1e3, 2.0 * m * N * K / ms / 1e9), flush=True)
print("encoder layer GEMMs: %.1f us -> %.1f TFLOP/s" % (tot_ms * 1e3, 2.0 * M * (1536 * 512 + 512 * 512 + 2 * 2048 * 512) / tot_ms / 1e9))
