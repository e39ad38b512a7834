"""Test infrastructure: the fp32 oracle (oracle/paraformer_ref.py + oracle/frontend.py) run over many segments in a pool of
worker processes, one intra-op thread each -- the way bench.py's cpu_baseline leg runs it -- so that the full-size model
(215.8 M parameters) can be checked against the CUDA path on dozens of segments in well under a minute.

Workers are SPAWNED (not forked): the calling pytest / bench process has usually initialised CUDA and OpenMP thread pools.
Every worker rebuilds the seeded synthetic weights itself, so nothing large crosses the process boundary on the way in."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_W = {}


def _init(cfg_over, seed, jitter_ln, transform):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import torch
    torch.set_num_threads(1)
    synth = importlib.import_module("asr-2pass_b200.synth")
    from oracle import paraformer_ref as R
    cfg, W = synth.make_weights(cfg_over, seed, jitter_ln)
    if transform:
        getattr(synth, transform)(W)
    means, vars_ = synth.make_cmvn(int(cfg["feat_dim"]))
    _W.update(W={k: torch.from_numpy(v) for k, v in W.items()}, pc=R.PfConfig.from_dict(cfg), means=means, vars=vars_)


def _one(args):
    idx, pcm16, want_logits = args
    from oracle import frontend as F
    from oracle import paraformer_ref as R
    x = pcm16.astype(np.float32) / np.float32(32768)
    feats = F.lfr_cmvn(F.fbank(x), _W["means"], _W["vars"])
    o = R.forward(feats, _W["W"], _W["pc"], want_taps=False)
    lg = o["logits"].numpy()
    top2 = np.sort(lg, axis=1)[:, -2:] if lg.shape[0] else np.zeros((0, 2), np.float32)
    out = dict(idx=idx, T=int(feats.shape[0]), enc=o["enc"].numpy(), alphas=o["alphas"].numpy(), fires=o["fires"].numpy(),
               token_num=int(o["token_num"]), ids=np.asarray(o["ids"], np.int32), top_gap=(top2[:, 1] - top2[:, 0]).astype(np.float32),
               logit_absmax=float(np.abs(lg).max()) if lg.size else 0.0)
    if want_logits:
        out["logits"] = lg
    return out


def oracle_forward_many(segments_pcm16, cfg_over=None, seed=0, jitter_ln=False, transform=None, procs=None, want_logits=True):
    """segments_pcm16: list of int16 arrays.  Returns the list of per-segment oracle outputs (dicts, see _one)."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    from oracle import frontend as F
    F.lib()   # build the C front end once, before the workers race for it
    P = procs or max(1, min(os.cpu_count() or 1, 32, len(segments_pcm16)))
    order = np.argsort([-len(s) for s in segments_pcm16])   # longest first: better balance
    with ProcessPoolExecutor(P, mp_context=mp.get_context("spawn"), initializer=_init, initargs=(cfg_over, seed, jitter_ln, transform)) as ex:
        res = list(ex.map(_one, [(int(i), segments_pcm16[int(i)], want_logits) for i in order]))
    out = [None] * len(segments_pcm16)
    for r in res:
        out[r["idx"]] = r
    return out
