"""Generates the committed golden vectors under tests/golden/ IN THE BUILD CONTAINER, where /root/reference
exists: front-end vectors come from the reference's own compiled kaldi-native-fbank (oracle/_ref, built by
oracle/Makefile from the sources where they lie); LFR/CMVN vectors from a literal transcription of
Paraformer::LfrCmvn's vector-insert algorithm (paraformer.cpp:421-461) written independently of the
oracle's index formula; model taps from the fp32 oracle (parity unpinned, see oracle/paraformer_ref.py).

    python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def lfr_cmvn_literal(feats, means, vars_, lfr_m=7, lfr_n=6):
    """Line-by-line transcription of Paraformer::LfrCmvn (vector inserts and all)."""
    asr_feats = [list(map(np.float32, f)) for f in feats]
    out_feats = []
    T = len(asr_feats)
    T_lrf = int(np.ceil(1.0 * T / lfr_n))
    for _ in range((lfr_m - 1) // 2):
        asr_feats.insert(0, list(asr_feats[0]))
    T = T + (lfr_m - 1) // 2
    for i in range(T_lrf):
        p = []
        if lfr_m <= T - i * lfr_n:
            for j in range(lfr_m):
                p.extend(asr_feats[i * lfr_n + j])
        else:
            num_padding = lfr_m - (T - i * lfr_n)
            for j in range(len(asr_feats) - i * lfr_n):
                p.extend(asr_feats[i * lfr_n + j])
            for _ in range(num_padding):
                p.extend(asr_feats[-1])
        out_feats.append(p)
    out = np.asarray(out_feats, np.float32).reshape(len(out_feats), -1)
    return ((out + means[None, :]).astype(np.float32) * vars_[None, :]).astype(np.float32)


def main():
    from oracle import frontend as F
    synth = importlib.import_module("asr-2pass_b200.synth")
    F.build(ref=True)
    assert F.ref_lib() is not None, "needs /root/reference (run in the build container)"
    means, vars_ = synth.make_cmvn()
    g = {}
    for n in (399, 400, 559, 560, 1359, 16000, 52800):
        pcm16 = synth.make_audio(n, 9000 + n)
        g["pcm_%d" % n] = pcm16
        fb = F.fbank_ref(pcm16.astype(np.float32) / np.float32(32768))
        g["fbank_%d" % n] = fb
    rng = np.random.default_rng(5)
    for n_fb in (1, 3, 6, 7, 12, 13, 98):
        fb = rng.standard_normal((n_fb, 80)).astype(np.float32)
        g["lfr_in_%d" % n_fb] = fb
        g["lfr_out_%d" % n_fb] = lfr_cmvn_literal(fb, means, vars_)
    # knf test-rfft.cc:32-50 known answer through the reference's Rfft
    import ctypes
    d = np.array([1, -1, 3, 8, 20, 6, 0, 2], np.float32)
    F.ref_lib().knf_ref_rfft(d.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 8)
    g["rfft8"] = d
    np.savez_compressed(os.path.join(HERE, "frontend_golden.npz"), **g)

    # small-architecture model taps from the fp32 oracle (2 encoder / 2 decoder layers, jittered LayerNorm)
    import torch
    from oracle import paraformer_ref as R
    cfg, W = synth.make_weights(dict(n_enc=2, n_dec=2), seed=0, jitter_ln=True)
    pc = R.PfConfig.from_dict(cfg)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    m = {}
    for n in (16000, 52800):
        pcm16 = g["pcm_%d" % n]
        feats = F.lfr_cmvn(g["fbank_%d" % n], means, vars_)
        o = R.forward(feats, Wt, pc)
        m["enc_%d" % n] = o["enc"].numpy().astype(np.float16)
        m["alphas_%d" % n] = o["alphas"].numpy()
        m["fires_%d" % n] = o["fires"].numpy()
        m["ids_%d" % n] = np.asarray(o["ids"], np.int32)
        m["token_num_%d" % n] = np.asarray([o["token_num"]], np.int32)
        lg = o["logits"].numpy()
        top2 = np.sort(lg, axis=1)[:, -2:]
        m["top_gap_%d" % n] = (top2[:, 1] - top2[:, 0]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "model_small_golden.npz"), **m)

    # config 3 (timestamp head + contextual decoder + hotword compiler), same small architecture, fp32 oracle
    cfg, W = synth.make_weights(dict(n_enc=2, n_dec=2, timestamp=1, contextual=1), seed=3, jitter_ln=True)
    pc = R.PfConfig.from_dict(cfg)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    rng = np.random.default_rng(77)
    hw_ids = np.zeros((9, 10), np.int32)
    hw_len = np.zeros(9, np.int32)
    for j in range(8):
        L = int(rng.integers(1, 8))
        hw_ids[j, :L] = rng.integers(3, pc.vocab - 1, L)
        hw_len[j] = L
    hw_ids[8, 0], hw_len[8] = 1, 1
    hw = R.select_hotword_rows(R.hotword_embed(hw_ids, Wt), hw_len)
    c3 = dict(hw_ids=hw_ids, hw_len=hw_len, hw_emb=hw.numpy())
    for n in (16000, 52800):
        feats = F.lfr_cmvn(g["fbank_%d" % n], means, vars_)
        o = R.forward(feats, Wt, pc, hw_emb=hw)
        c3["us_alphas_%d" % n] = o["us_alphas"].numpy()
        c3["us_peaks_%d" % n] = o["us_peaks"].numpy()
        c3["ids_%d" % n] = np.asarray(o["ids"], np.int32)
        c3["token_num_%d" % n] = np.asarray([o["token_num"]], np.int32)
        lg = o["logits"].numpy()
        top2 = np.sort(lg, axis=1)[:, -2:]
        c3["top_gap_%d" % n] = (top2[:, 1] - top2[:, 0]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "model_cfg3_golden.npz"), **c3)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
