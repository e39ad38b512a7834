import importlib, os, sys
sys.path.insert(0, '/root/repo')
capi = importlib.import_module("asr-2pass_b200.capi")
M = 61440
for name, N, K, mode in [("out+mem+res", 512, 512, 3), ("ffn2 res", 512, 2048, 2), ("qkv", 1536, 512, 0)]:
    ms = capi.op_gemm_bench(M, N, K, mode, 20)
    print("%-14s dbg=%s %8.1f us %7.1f TF" % (name, os.environ.get("B200PF_GEMM_DBG", "0"), ms * 1e3, 2.0 * M * N * K / ms / 1e9), flush=True)
