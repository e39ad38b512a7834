import importlib, os, sys, tempfile
import numpy as np
sys.path.insert(0, '/root/repo')
capi = importlib.import_module("asr-2pass_b200.capi")
synth = importlib.import_module("asr-2pass_b200.synth")
tmp = tempfile.mkdtemp(prefix="b200pf_tr_")
synth.write_synthetic_model_dir(tmp, None, seed=0)
n = capi.device_count()
h = capi.OfflineHandle(tmp, max_rows=65536, max_segments=4096, batch_size=4096, devices=list(range(n)))
one = synth.make_audio(960000, 4321)
pcm5 = np.tile(one, 256)
b5 = [i * 960000 for i in range(256)]
e5 = [(i + 1) * 960000 for i in range(256)]
for _ in range(2):
    h.infer_segments(pcm5, b5, e5)
os.environ["X"] = "1"
print("=== traced call", file=sys.stderr, flush=True)
import time
t0 = time.perf_counter()
h.infer_segments(pcm5, b5, e5)
print("python wall %.2f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
