#!/usr/bin/env python
"""Runs the reference-style benchmark harness (tools/funasr_b200_offline_rtf.cpp = the reference's
onnxruntime/bin/funasr-onnx-offline-rtf.cpp pattern) on synthetic wavs: P threads share one FunOfflineInit handle, each
FunOfflineInfer call decodes one wav.  Twice: calls as they come (every wav is its own batch-1 forward), and with the
shim's micro-batcher merging the concurrent calls.  Prints the harness' JSON lines.

    python tools/run_rtf_harness.py [--wavs 256] [--threads 64] [--devices 0]
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--wavs", type=int, default=256)
    ap.add_argument("--threads", type=int, default=64)
    ap.add_argument("--devices", default="0")
    args = ap.parse_args()
    synth = importlib.import_module("asr-2pass_b200.synth")
    exe = os.path.join(ROOT, "asr-2pass_b200", "lib", "funasr-b200-offline-rtf")
    tmp = tempfile.mkdtemp(prefix="b200pf_rtf_")
    model = os.path.join(tmp, "model")
    synth.write_synthetic_model_dir(model, None, seed=0)
    pcm, offs = synth.make_segments(1024)
    pick = np.linspace(0, 1023, args.wavs).astype(int)
    scp = os.path.join(tmp, "wav.scp")
    with open(scp, "w") as f:
        for k, i in enumerate(pick):
            path = os.path.join(tmp, "utt%04d.wav" % k)
            with wave.open(path, "wb") as w:
                w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
                w.writeframes(pcm[offs[i]:offs[i + 1]].astype("<i2").tobytes())
            f.write("utt%04d %s\n" % (k, path))
    out = []
    for mb in ("0", "20000"):
        cmd = [exe, "--model-dir", model, "--wav-scp", scp, "--thread-num", str(args.threads), "--devices", args.devices, "--max-rows", "65536"]
        if mb != "0":
            cmd += ["--micro-batch-us", mb]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            print(r.stdout[-2000:], r.stderr[-2000:], file=sys.stderr)
            raise SystemExit("harness failed")
        out.append(json.loads(lines[-1]))
        print(lines[-1], flush=True)
    return out


if __name__ == "__main__":
    main()
