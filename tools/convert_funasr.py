#!/usr/bin/env python
"""Writes the flat weight file the B200 engines read from a FunASR PyTorch checkpoint (a state_dict saved with torch.save, possibly
wrapped as {"state_dict": ...} / {"model": ...}) and, when given, the model's config.yaml.  Tensor names are kept as they are
(upstream FunASR names); dimensions are read off the tensor shapes, everything the shapes cannot tell from the yaml / the flags.

    python tools/convert_funasr.py am   --checkpoint model.pt [--config config.yaml] --out-dir <model-dir>   -> model.b200pf
    python tools/convert_funasr.py vad  --checkpoint model.pt --out-dir <vad-dir>                            -> vad.b200pf
    python tools/convert_funasr.py punc --checkpoint model.pt [--config config.yaml] --out-dir <punc-dir>    -> punc.b200pf, punc_list.json

The reference's other files (am.mvn, tokens.json, config.yaml, seg_dict) are used as they are; put them in the same directory.
No real checkpoint is available offline, so this tool is exercised on synthetic state_dicts only (tests/test_host_cpu.py):
PARITY WITH REAL EXPORTED WEIGHTS IS UNVERIFIED.  The acoustic network's restatement (oracle/paraformer_ref.py) is itself unpinned;
hyper-parameters the kernels hard-code are checked against config.yaml (check_supported) and a mismatch stops the conversion."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_state_dict(path):
    import torch
    sd = torch.load(path, map_location="cpu", weights_only=True)
    for key in ("state_dict", "model", "model_state_dict"):
        if isinstance(sd, dict) and key in sd and isinstance(sd[key], dict):
            sd = sd[key]
    return {k: v.detach().float().numpy() for k, v in sd.items() if hasattr(v, "detach")}


def load_yaml(path):
    if not path:
        return {}
    import yaml
    with open(path, encoding="utf-8") as f:
        return yaml.safe_load(f) or {}


def count(sd, prefix, suffix):
    n = 0
    while "%s%d%s" % (prefix, n, suffix) in sd:
        n += 1
    return n


def check_supported(y):
    """The engine's kernels fix some of the architecture's hyper-parameters (elementwise.cu cif_alpha_kernel: smooth_factor 1.0 /
    noise_threshold 0.0; fsmn_kernel: symmetric taps, sanm_shfit 0; the timestamp head reads the encoder output, i.e.
    use_cif1_cnn false with upsample_type cnn_blstm; ln_eps 1e-12 and the V2 predictor without the +enc residual are written into
    the file).  A checkpoint trained with other values would convert and then produce wrong text, so every such key that
    config.yaml carries is checked here and the conversion stops on the first value the engine does not implement."""
    enc, pred, dec = y.get("encoder_conf", {}) or {}, y.get("predictor_conf", {}) or {}, y.get("decoder_conf", {}) or {}
    want = [
        ("encoder_conf.sanm_shfit", enc.get("sanm_shfit", enc.get("sanm_shift")), 0),
        ("encoder_conf.normalize_before", enc.get("normalize_before"), True),
        ("encoder_conf.input_layer", enc.get("input_layer"), "pe"),
        ("encoder_conf.selfattention_layer_type", enc.get("selfattention_layer_type"), "sanm"),
        ("decoder_conf.sanm_shfit", dec.get("sanm_shfit", dec.get("sanm_shift")), 0),
        ("predictor_conf.smooth_factor", pred.get("smooth_factor"), 1.0),
        ("predictor_conf.noise_threshold", pred.get("noise_threshold"), 0.0),
        ("predictor_conf.l_order", pred.get("l_order"), 1),
        ("predictor_conf.r_order", pred.get("r_order"), 1),
        ("predictor_conf.use_cif1_cnn", pred.get("use_cif1_cnn"), False),
        ("predictor_conf.upsample_type", pred.get("upsample_type"), "cnn_blstm"),
        ("predictor_conf.upsample_times", pred.get("upsample_times"), 3),
        ("frontend_conf.lfr_m", (y.get("frontend_conf", {}) or {}).get("lfr_m"), 7),
        ("frontend_conf.lfr_n", (y.get("frontend_conf", {}) or {}).get("lfr_n"), 6),
        ("frontend_conf.n_mels", (y.get("frontend_conf", {}) or {}).get("n_mels"), 80),
        ("frontend_conf.window", (y.get("frontend_conf", {}) or {}).get("window"), "hamming"),
    ]
    for key, got, need in want:
        if got is None:
            continue                       # not in this config.yaml: the upstream default is the supported value
        same = (abs(float(got) - float(need)) < 1e-9) if isinstance(need, float) else (got == need)
        if not same:
            raise SystemExit("unsupported %s = %r: the B200 engine implements %r only (tools/convert_funasr.py check_supported)" % (key, got, need))
    name = str(y.get("predictor", "") or "")
    if name and "cif" not in name.lower():
        raise SystemExit("unsupported predictor %r" % name)


def am_config(sd, y, n_heads):
    check_supported(y)
    enc, pred, dec = y.get("encoder_conf", {}) or {}, y.get("predictor_conf", {}) or {}, y.get("decoder_conf", {}) or {}
    cfg = dict(
        feat_dim=sd["encoder.encoders0.0.norm1.weight"].shape[0], d_model=sd["encoder.after_norm.weight"].shape[0],
        n_heads=int(enc.get("attention_heads", n_heads)), d_ff=sd["encoder.encoders0.0.feed_forward.w_1.weight"].shape[0],
        n_enc=1 + count(sd, "encoder.encoders.", ".norm1.weight"), kernel=sd["encoder.encoders0.0.self_attn.fsmn_block.weight"].shape[-1],
        vocab=sd["decoder.output_layer.weight"].shape[0], cif_threshold=float(pred.get("threshold", 1.0)),
        tail_threshold=float(pred.get("tail_threshold", 0.45)), pred_residual=0, ln_eps=1e-12)
    cfg["contextual"] = int("bias_embed.weight" in sd)
    cfg["n_dec"] = count(sd, "decoder.decoders.", ".norm1.weight") + cfg["contextual"]
    cfg["timestamp"] = int("predictor.upsample_cnn.weight" in sd)
    if cfg["timestamp"]:
        cfg["us_times"] = sd["predictor.upsample_cnn.weight"].shape[-1]
        cfg["smooth_factor2"] = float(pred.get("smooth_factor2", 0.25))
        cfg["noise_threshold2"] = float(pred.get("noise_threshold2", 0.01))
    if int(dec.get("att_layer_num", cfg["n_dec"])) != cfg["n_dec"]:
        raise SystemExit("decoder_conf.att_layer_num disagrees with the checkpoint")
    return cfg


def punc_config(sd, y, n_heads):
    enc = y.get("encoder_conf", {}) or {}
    return dict(vocab=sd["embed.weight"].shape[0], d_model=sd["embed.weight"].shape[1], n_heads=int(enc.get("attention_heads", n_heads)),
                d_ff=sd["encoder.encoders0.0.feed_forward.w_1.weight"].shape[0], n_layers=1 + count(sd, "encoder.encoders.", ".norm1.weight"),
                kernel=sd["encoder.encoders0.0.self_attn.fsmn_block.weight"].shape[-1], n_punc=sd["decoder.weight"].shape[0], ln_eps=1e-12,
                sanm_shift=int(enc.get("sanm_shfit", enc.get("sanm_shift", 0))))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("kind", choices=["am", "vad", "punc"])
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--config", default=None)
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--n-heads", type=int, default=None, help="attention heads when config.yaml is not given (am: 4, punc: 8)")
    args = ap.parse_args(argv)
    mf = importlib.import_module("asr-2pass_b200.modelfile")
    synth = importlib.import_module("asr-2pass_b200.synth")
    sd, y = load_state_dict(args.checkpoint), load_yaml(args.config)
    os.makedirs(args.out_dir, exist_ok=True)
    if args.kind == "am":
        cfg = am_config(sd, y, args.n_heads or 4)
        want = synth.param_shapes(cfg)
        name = "model.b200pf"
    elif args.kind == "vad":
        cfg = dict(synth.VAD_DIMS)
        want = synth.vad_param_shapes()
        name = "vad.b200pf"
    else:
        cfg = punc_config(sd, y, args.n_heads or 8)
        want = synth.punc_param_shapes(cfg)
        name = "punc.b200pf"
        punc_list = ((y.get("model_conf", {}) or {}).get("punc_list")) or synth.PUNC_LIST
        with open(os.path.join(args.out_dir, "punc_list.json"), "w", encoding="utf-8") as f:
            json.dump([str(p) for p in punc_list], f, ensure_ascii=False)
    missing = [k for k in want if k not in sd]
    if missing:
        raise SystemExit("checkpoint lacks %d tensors the engine needs, e.g. %s" % (len(missing), missing[:4]))
    bad = [k for k, shp in want.items() if tuple(sd[k].shape) != tuple(shp)]
    if bad:
        raise SystemExit("unexpected shapes, e.g. %s: %s, expected %s" % (bad[0], sd[bad[0]].shape, want[bad[0]]))
    mf.write_weights(os.path.join(args.out_dir, name), cfg, {k: sd[k] for k in want})
    print(json.dumps(dict(wrote=os.path.join(args.out_dir, name), config=cfg, tensors=len(want), ignored=len(sd) - len(want))))


if __name__ == "__main__":
    main()
