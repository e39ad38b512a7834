#!/usr/bin/env python
"""Punctuation throughput on one B200: N transcripts of ~1 h of speech each (about 15 k tokens),
full-size random-init CT-Transformer (vocab 272727, d 256, 4 layers), through funasr_b200::CTTransformerB200.

  * one request at a time (AddPunc: the reference's shape, one network call per 20-token mini-sentence);
  * the same requests in lock step (AddPuncBatch: one network call per round for all of them);
  * the oracle's fp32 CPU restatement of the same walk on a bounded sample, as the CPU baseline ("port").

Not a bench.py line (bench.py measures configs[1]); the JSON it prints is kept under profiles/.
    python tools/bench_punc.py [--requests 64] [--tokens 15000]
"""
import argparse
import importlib
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--requests", type=int, default=64)
    ap.add_argument("--tokens", type=int, default=15000)
    ap.add_argument("--cpu-tokens", type=int, default=600)
    args = ap.parse_args()
    import torch
    synth = importlib.import_module("asr-2pass_b200.synth")
    capi = importlib.import_module("asr-2pass_b200.capi")
    from oracle import punc_ref as PR
    d = tempfile.mkdtemp(prefix="b200pf_punc_")
    cfg, W, toks = synth.write_synthetic_punc_dir(d, None, seed=0)
    host = capi.HostPunc(d, max_tokens=65536)
    tok = PR.Tokenizer(toks)
    texts = [synth.make_text(args.tokens, 500 + i, toks) for i in range(args.requests)]
    n_tok = [len(tok.tokenize(t)[1]) for t in texts]
    host.add_punc(texts[0][:2000])                       # warm-up
    t0 = time.perf_counter()
    single = [host.add_punc(t) for t in texts[:4]]
    t_single = (time.perf_counter() - t0) / 4
    host.add_punc_batch(texts[:2])
    t0 = time.perf_counter()
    batch, rounds = host.add_punc_batch(texts)
    t_batch = time.perf_counter() - t0
    assert batch[:4] == single
    # CPU baseline: the oracle's walk + fp32 network on a bounded sample
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    sample = synth.make_text(args.cpu_tokens, 499, toks)
    t0 = time.perf_counter()
    PR.add_punc(sample, tok, lambda ids: PR.infer_ids(PR.forward(ids, Wt, cfg).numpy()))
    t_cpu = time.perf_counter() - t0
    n_cpu = len(tok.tokenize(sample)[1])
    out = dict(
        config="CT-Transformer vocab 272727 d256 h8 ff1024 L4 (random init); %d requests x ~%d tokens (a 1 h transcript each)" % (args.requests, args.tokens),
        tokens_total=int(sum(n_tok)), rounds=rounds,
        single_request_ms=t_single * 1e3, single_request_tokens_per_s=n_tok[0] / t_single,
        batch_ms=t_batch * 1e3, batch_tokens_per_s=sum(n_tok) / t_batch, batch_requests_per_s=args.requests / t_batch,
        cpu_port_tokens_per_s=n_cpu / t_cpu, cpu_sample="%d tokens, torch fp32, %d threads" % (n_cpu, torch.get_num_threads()),
        identical_results=True)
    print(json.dumps(out))
    host.close()


if __name__ == "__main__":
    main()
