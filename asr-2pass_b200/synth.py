"""Synthetic workloads and random-init model directories (SURVEY.md §8(d) "Weights" and configs 1-5).

There is no network: real checkpoints and speech corpora are unavailable, so both sides of every parity
test and every benchmark line use these seeded generators.  Nothing here is on the product hot path.
"""
import math

import numpy as np

# PfConfig defaults (mirrors csrc/config.h)
DEFAULT_CFG = dict(feat_dim=560, d_model=512, n_heads=4, d_ff=2048, n_enc=50, n_dec=16, kernel=11,
                   vocab=8404, cif_threshold=1.0, tail_threshold=0.45, pred_residual=0, ln_eps=1e-12,
                   timestamp=0, contextual=0, us_times=3, smooth_factor2=0.25, noise_threshold2=0.01)


def param_shapes(cfg):
    D, Fd, V, K = int(cfg["d_model"]), int(cfg["d_ff"]), int(cfg["vocab"]), int(cfg["kernel"])
    out = {}

    def ln(p, n):
        out[p + ".weight"] = (n,)
        out[p + ".bias"] = (n,)

    def lin(p, o, i, bias=True):
        out[p + ".weight"] = (o, i)
        if bias:
            out[p + ".bias"] = (o,)

    for l in range(int(cfg["n_enc"])):
        p = "encoder.encoders0.0" if l == 0 else "encoder.encoders.%d" % (l - 1)
        din = int(cfg["feat_dim"]) if l == 0 else D
        ln(p + ".norm1", din)
        lin(p + ".self_attn.linear_q_k_v", 3 * D, din)
        out[p + ".self_attn.fsmn_block.weight"] = (D, 1, K)
        lin(p + ".self_attn.linear_out", D, D)
        ln(p + ".norm2", D)
        lin(p + ".feed_forward.w_1", Fd, D)
        lin(p + ".feed_forward.w_2", D, Fd)
    ln("encoder.after_norm", D)
    out["predictor.cif_conv1d.weight"] = (D, D, 3)
    out["predictor.cif_conv1d.bias"] = (D,)
    lin("predictor.cif_output", 1, D)
    ctx = int(cfg.get("contextual", 0))
    if int(cfg.get("timestamp", 0)):   # CifPredictorV3 upsample head (SURVEY.md Appendix B, config-3 extras)
        out["predictor.upsample_cnn.weight"] = (D, D, int(cfg.get("us_times", 3)))
        out["predictor.upsample_cnn.bias"] = (D,)
        for sfx in ("", "_reverse"):
            out["predictor.blstm.weight_ih_l0" + sfx] = (4 * D, D)
            out["predictor.blstm.weight_hh_l0" + sfx] = (4 * D, D)
            out["predictor.blstm.bias_ih_l0" + sfx] = (4 * D,)
            out["predictor.blstm.bias_hh_l0" + sfx] = (4 * D,)
        lin("predictor.cif_output2", 1, 2 * D)
    dec_names = ["decoder.decoders.%d" % l for l in range(int(cfg["n_dec"]) - (1 if ctx else 0))]
    if ctx:                            # ContextualParaformerDecoder + hotword compiler (model_eb)
        dec_names.append("decoder.last_decoder")
        ln("decoder.bias_decoder.norm3", D)
        lin("decoder.bias_decoder.src_attn.linear_q", D, D)
        lin("decoder.bias_decoder.src_attn.linear_k_v", 2 * D, D)
        lin("decoder.bias_decoder.src_attn.linear_out", D, D)
        out["decoder.bias_output.weight"] = (D, 2 * D, 1)
        out["bias_embed.weight"] = (V, D)
        out["bias_encoder.weight_ih_l0"] = (4 * D, D)
        out["bias_encoder.weight_hh_l0"] = (4 * D, D)
        out["bias_encoder.bias_ih_l0"] = (4 * D,)
        out["bias_encoder.bias_hh_l0"] = (4 * D,)
    for p in dec_names:
        ln(p + ".norm1", D)
        lin(p + ".feed_forward.w_1", Fd, D)
        ln(p + ".feed_forward.norm", Fd)
        lin(p + ".feed_forward.w_2", D, Fd, bias=False)
        ln(p + ".norm2", D)
        out[p + ".self_attn.fsmn_block.weight"] = (D, 1, K)
        ln(p + ".norm3", D)
        lin(p + ".src_attn.linear_q", D, D)
        lin(p + ".src_attn.linear_k_v", 2 * D, D)
        lin(p + ".src_attn.linear_out", D, D)
    p = "decoder.decoders3.0"
    ln(p + ".norm1", D)
    lin(p + ".feed_forward.w_1", Fd, D)
    ln(p + ".feed_forward.norm", Fd)
    lin(p + ".feed_forward.w_2", D, Fd, bias=False)
    ln("decoder.after_norm", D)
    lin("decoder.output_layer", V, D)
    return out


def make_weights(cfg=None, seed=0, jitter_ln=False):
    """nn.Linear / nn.Conv1d default init U(+-1/sqrt(fan_in)) for weights and biases; LayerNorm
    gamma=1, beta=0 (or jittered, to exercise gamma/beta in tests); cif_output.bias = 0."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    rng = np.random.default_rng(seed)
    W = {}
    shapes = param_shapes(cfg)
    for name, shp in shapes.items():
        is_ln = (".norm" in name) or name.endswith("after_norm.weight") or name.endswith("after_norm.bias")
        if is_ln:
            if name.endswith(".weight"):
                W[name] = (1.0 + (0.1 * rng.uniform(-1, 1, shp) if jitter_ln else 0.0) * np.ones(shp)).astype(np.float32)
            else:
                W[name] = ((0.1 * rng.uniform(-1, 1, shp)) if jitter_ln else np.zeros(shp)).astype(np.float32)
            continue
        if name == "bias_embed.weight":      # nn.Embedding
            W[name] = rng.uniform(-1, 1, shp).astype(np.float32)
        elif "weight_ih_l0" in name or "weight_hh_l0" in name or "bias_ih_l0" in name or "bias_hh_l0" in name:
            W[name] = rng.uniform(-1, 1, shp).astype(np.float32) / np.float32(math.sqrt(int(cfg["d_model"])))  # nn.LSTM: 1/sqrt(hidden)
        elif name.endswith(".weight"):
            fan_in = int(np.prod(shp[1:]))
            W[name] = rng.uniform(-1, 1, shp).astype(np.float32) / np.float32(math.sqrt(fan_in))
        else:
            wshape = shapes[name[:-5] + ".weight"]
            fan_in = int(np.prod(wshape[1:]))
            W[name] = rng.uniform(-1, 1, shp).astype(np.float32) / np.float32(math.sqrt(fan_in))
    W["predictor.cif_output.bias"] = np.zeros((1,), np.float32)
    return cfg, W


def make_peaky(W, live=64):
    """In place: a "peaky" variant of the random-init model for exact-equality tests.  Only `live` vocabulary entries stay
    reachable (the bias of every other class drops by 30), so the top-1 margin of a row is that of the best of 64 Gaussians
    instead of 8404 -- several times the logit rounding error -- and segments whose every row and every CIF fire clears its margin
    can be picked (tests/golden/peaky_cases.json).  Architecture, shapes and every other tensor are unchanged."""
    b = W["decoder.output_layer.bias"]
    b[live:] -= np.float32(30.0)
    return W


VAD_DIMS = dict(input_dim=400, input_affine_dim=140, linear_dim=250, proj_dim=128, lorder=20, n_layers=4, output_affine_dim=140, output_dim=248)


def vad_param_shapes(d=VAD_DIMS):
    """FSMN-VAD (FunASR fsmn_vad_streaming `FSMN`) parameter names and shapes."""
    out = {}
    out["encoder.in_linear1.linear.weight"] = (d["input_affine_dim"], d["input_dim"])
    out["encoder.in_linear1.linear.bias"] = (d["input_affine_dim"],)
    out["encoder.in_linear2.linear.weight"] = (d["linear_dim"], d["input_affine_dim"])
    out["encoder.in_linear2.linear.bias"] = (d["linear_dim"],)
    for i in range(d["n_layers"]):
        p = "encoder.fsmn.%d" % i
        out[p + ".linear.linear.weight"] = (d["proj_dim"], d["linear_dim"])
        out[p + ".fsmn_block.conv_left.weight"] = (d["proj_dim"], 1, d["lorder"], 1)
        out[p + ".affine.linear.weight"] = (d["linear_dim"], d["proj_dim"])
        out[p + ".affine.linear.bias"] = (d["linear_dim"],)
    out["encoder.out_linear1.linear.weight"] = (d["output_affine_dim"], d["linear_dim"])
    out["encoder.out_linear1.linear.bias"] = (d["output_affine_dim"],)
    out["encoder.out_linear2.linear.weight"] = (d["output_dim"], d["output_affine_dim"])
    out["encoder.out_linear2.linear.bias"] = (d["output_dim"],)
    return out


def make_vad_weights(seed=0):
    rng = np.random.default_rng(seed)
    W = {}
    shapes = vad_param_shapes()
    for name, shp in shapes.items():
        fan_in = int(np.prod(shp[1:])) if name.endswith(".weight") else int(np.prod(shapes[name[:-5] + ".weight"][1:]))
        gain = 2.45 if name.endswith(".weight") and "conv_left" not in name else 1.0   # He-like: keeps activations O(1) through the ReLUs
        W[name] = (rng.uniform(-1, 1, shp) * gain).astype(np.float32) / np.float32(math.sqrt(fan_in))
    # a posterior a VAD could plausibly produce: moderate logits, and pdf 0 (silence) competing with the other 247 classes so
    # that the silence probability really moves between frames
    W["encoder.out_linear2.linear.weight"] *= np.float32(0.05)
    W["encoder.out_linear2.linear.weight"][0] *= np.float32(6.0)
    W["encoder.out_linear2.linear.bias"][0] += np.float32(5.5)
    return W


def write_synthetic_vad_dir(path, seed=0):
    """<vad-dir>/{am.mvn (400-dim), vad.b200pf}: the reference's VAD directory layout (com-define.h:56-58) with the flat weight
    file in place of model.onnx."""
    import os
    from . import modelfile
    W = make_vad_weights(seed)
    means, vars_ = make_cmvn(400)
    os.makedirs(path, exist_ok=True)
    modelfile.write_weights(os.path.join(path, "vad.b200pf"), dict(VAD_DIMS), W)
    modelfile.write_am_mvn(os.path.join(path, "am.mvn"), means, vars_)
    return W, means, vars_


def make_cmvn(n=560):
    j = np.arange(n, dtype=np.float64)
    means = (-(8.0 + 2.0 * np.sin(j))).astype(np.float32)
    vars_ = (0.15 + 0.05 * np.cos(j)).astype(np.float32)
    return means, vars_


def make_tokens(vocab=8404):
    toks = ["<blank>", "<s>", "</s>"]
    n_latin = 500 if vocab >= 2000 else max(4, vocab // 8)
    n_cjk = vocab - 4 - n_latin
    toks += [chr(0x4E00 + i) for i in range(n_cjk)]
    latin = [chr(ord("a") + i) for i in range(26)]
    i = 0
    while len(latin) < n_latin:
        a, b, c = i % 26, (i // 26) % 26, (i // 676) % 26
        w = chr(97 + a) + chr(97 + b) + (chr(97 + c) if i % 3 == 0 else "")
        latin.append(w + "@@" if i % 2 == 0 else w)
        i += 1
    # keep order but drop accidental duplicates deterministically
    seen, uniq = set(), []
    for w in latin:
        while w in seen:
            w = "x" + w
        seen.add(w)
        uniq.append(w)
    toks += uniq[:n_latin]
    toks.append("<unk>")
    assert len(toks) == vocab, (len(toks), vocab)
    return toks


def make_audio(n_samples, seed, as_int16=True):
    """Speech-like noise: white noise -> 2nd-order band-pass 100-4000 Hz -> 4 Hz raised-cosine syllable
    envelope, peak 0.3 full scale."""
    from scipy.signal import butter, lfilter
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n_samples)
    b, a = butter(1, [100.0, 4000.0], btype="bandpass", fs=16000.0)
    x = lfilter(b, a, x)
    t = np.arange(n_samples) / 16000.0
    env = 0.5 * (1.0 - np.cos(2.0 * np.pi * 4.0 * t + rng.uniform(0, 2 * np.pi)))
    x = x * (0.05 + 0.95 * env)
    x = 0.3 * x / (np.abs(x).max() + 1e-12)
    s = np.round(x * 32768.0).clip(-32768, 32767).astype(np.int16)
    return s if as_int16 else (s.astype(np.float32) / np.float32(32768.0))


def segment_lengths(n_segments=1024, lo_s=2.0, hi_s=20.0, seed=20260101):
    """Durations U[lo,hi] s quantised to 10 ms -> sample counts (config 2)."""
    rng = np.random.default_rng(seed)
    dur = np.round(rng.uniform(lo_s, hi_s, n_segments) * 100.0) / 100.0
    return (dur * 16000.0 + 0.5).astype(np.int64)


def make_segments(n_segments=1024, lo_s=2.0, hi_s=20.0, seed=20260101, audio_seed=1234):
    """Returns (pcm int16 concatenated, offsets int64 [n+1])."""
    lens = segment_lengths(n_segments, lo_s, hi_s, seed)
    offs = np.zeros(n_segments + 1, np.int64)
    offs[1:] = np.cumsum(lens)
    # one long generation pass, then per-segment envelopes are already varied by phase; re-seed per
    # 64-segment group to stay O(minutes) -> O(seconds) at 1024 segments
    pcm = np.empty(int(offs[-1]), np.int16)
    g = 64
    for s in range(0, n_segments, g):
        e = min(n_segments, s + g)
        pcm[offs[s]:offs[e]] = make_audio(int(offs[e] - offs[s]), audio_seed + s)
    return pcm, offs


def write_synthetic_model_dir(path, cfg=None, seed=0, jitter_ln=False, lang="zh-cn"):
    from . import modelfile
    cfg, W = make_weights(cfg, seed, jitter_ln)
    means, vars_ = make_cmvn(int(cfg["feat_dim"]))
    toks = make_tokens(int(cfg["vocab"]))
    modelfile.write_model_dir(path, cfg, W, means, vars_, toks, lang=lang)
    return cfg, W, means, vars_, toks


# ---- CT-Transformer punctuation model ---------------------------------------------------------
PUNC_CFG = dict(vocab=272727, d_model=256, n_heads=8, d_ff=1024, n_layers=4, kernel=11, n_punc=6, ln_eps=1e-12, sanm_shift=0)
PUNC_LIST = ["<unk>", "_", "，", "。", "？", "、"]


def punc_param_shapes(cfg):
    D, Fd, K = int(cfg["d_model"]), int(cfg["d_ff"]), int(cfg["kernel"])
    out = {"embed.weight": (int(cfg["vocab"]), D)}
    for l in range(int(cfg["n_layers"])):
        p = "encoder.encoders0.0" if l == 0 else "encoder.encoders.%d" % (l - 1)
        for nm, shp in ((".norm1.weight", (D,)), (".norm1.bias", (D,)), (".self_attn.linear_q_k_v.weight", (3 * D, D)),
                        (".self_attn.linear_q_k_v.bias", (3 * D,)), (".self_attn.fsmn_block.weight", (D, 1, K)),
                        (".self_attn.linear_out.weight", (D, D)), (".self_attn.linear_out.bias", (D,)), (".norm2.weight", (D,)),
                        (".norm2.bias", (D,)), (".feed_forward.w_1.weight", (Fd, D)), (".feed_forward.w_1.bias", (Fd,)),
                        (".feed_forward.w_2.weight", (D, Fd)), (".feed_forward.w_2.bias", (D,))):
            out[p + nm] = shp
    out["encoder.after_norm.weight"] = (D,)
    out["encoder.after_norm.bias"] = (D,)
    out["decoder.weight"] = (int(cfg["n_punc"]), D)
    out["decoder.bias"] = (int(cfg["n_punc"]),)
    return out


def make_punc_weights(cfg=None, seed=0):
    """Default torch inits (Embedding N(0,1), Linear / Conv1d U(+-1/sqrt(fan_in)), LayerNorm jittered); the classifier is scaled up
    and biased so that the five reachable classes all occur (a punctuation model that never fires would not exercise AddPunc)."""
    cfg = dict(PUNC_CFG, **(cfg or {}))
    rng = np.random.default_rng(seed)
    W = {}
    shapes = punc_param_shapes(cfg)
    for name, shp in shapes.items():
        if name == "embed.weight":
            W[name] = rng.standard_normal(shp).astype(np.float32)
        elif ".norm" in name or "after_norm" in name:
            W[name] = ((1.0 if name.endswith(".weight") else 0.0) + 0.1 * rng.uniform(-1, 1, shp)).astype(np.float32)
        elif name.endswith(".weight"):
            W[name] = (rng.uniform(-1, 1, shp) / math.sqrt(int(np.prod(shp[1:])))).astype(np.float32)
        else:
            wshape = shapes[name[:-5] + ".weight"]
            W[name] = (rng.uniform(-1, 1, shp) / math.sqrt(int(np.prod(wshape[1:])))).astype(np.float32)
    W["decoder.weight"] *= np.float32(3.0)
    W["decoder.bias"] = np.asarray([-1.5, 1.8, 0.6, 0.0, -0.6, 0.0][:int(cfg["n_punc"])], np.float32)
    return cfg, W


def make_punc_tokens(vocab):
    """<unk> first (the reference looks it up by name), CJK from U+4E00, lower-case latin words, digits."""
    toks = ["<unk>"]
    n_latin = min(2000, max(8, vocab // 8))
    toks += [chr(0x4E00 + i) for i in range(min(20000, vocab - 1 - n_latin))]
    i = 0
    while len(toks) < vocab:                 # bijective base-26 words: a .. z, aa, ab, ...
        n, w = i, ""
        while True:
            w = chr(97 + n % 26) + w
            n = n // 26 - 1
            if n < 0:
                break
        toks.append(w)
        i += 1
    assert len(toks) == vocab and len(set(toks)) == vocab
    return toks


def write_synthetic_punc_dir(path, cfg=None, seed=0):
    """<punc-dir>/{punc.b200pf, tokens.json, punc_list.json, config.yaml}: config.yaml carries model_conf.punc_list the way the
    reference's CTokenizer::OpenYaml reads it (tokenizer.cpp:136-160); punc_list.json is the same list for the B200 host side."""
    import json
    import os
    from . import modelfile
    cfg, W = make_punc_weights(cfg, seed)
    toks = make_punc_tokens(int(cfg["vocab"]))
    os.makedirs(path, exist_ok=True)
    modelfile.write_weights(os.path.join(path, "punc.b200pf"), cfg, W)
    with open(os.path.join(path, "tokens.json"), "w", encoding="utf-8") as f:
        json.dump(toks, f, ensure_ascii=False)
    with open(os.path.join(path, "punc_list.json"), "w", encoding="utf-8") as f:
        json.dump(PUNC_LIST, f, ensure_ascii=False)
    with open(os.path.join(path, "config.yaml"), "w", encoding="utf-8") as f:
        f.write("model: CTTransformer\nmodel_conf:\n  punc_list:\n" + "".join('  - "%s"\n' % p for p in PUNC_LIST))
    return cfg, W, toks


def make_text(n_tokens, seed, toks, english_frac=0.15):
    """Unpunctuated transcript-like text over the punctuation vocabulary: runs of CJK characters, lower/upper-case latin words
    separated by spaces, the odd out-of-vocabulary character."""
    rng = np.random.default_rng(seed)
    cjk = [t for t in toks[1:4000] if ord(t[0]) >= 0x4E00]
    lat = [t for t in toks if t and ord(t[0]) < 128 and t != "<unk>"][:500]
    out, n = [], 0
    while n < n_tokens:
        if rng.random() < english_frac and lat:
            k = int(rng.integers(1, 4))
            ws = [lat[int(rng.integers(len(lat)))] for _ in range(k)]
            if rng.random() < 0.2:
                ws[0] = ws[0].upper()
            out.append(" " + " ".join(ws) + " ")
            n += k
        else:
            k = int(rng.integers(1, 12))
            s = "".join(cjk[int(rng.integers(len(cjk)))] for _ in range(k))
            if rng.random() < 0.05:
                s += "㐀"            # not in the vocabulary -> <unk>
            out.append(s)
            n += len(s)
    return "".join(out).strip()
