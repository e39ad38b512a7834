import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
capi = importlib.import_module("asr-2pass_b200.capi")
T = np.array([int(x) for x in sys.argv[1:]] or [33], np.int32)
print(capi.op_attention_bench(T, iters=1), flush=True)
