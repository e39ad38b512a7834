"""ORACLE (test infrastructure only) — fp32 CPU restatement of the Paraformer-large graph that the
reference executes through onnxruntime (`Ort::Session::Run`, onnxruntime/src/paraformer.cpp:541).

The arithmetic lives in a third-party dependency that is NOT under /root/reference: the graph is the
FunASR export of `speech_paraformer-large_asr_nat-zh-cn-16k-common-vocab8404` rev v2.0.5
(funasr==0.7.5, dj_py310_environment.yml:153; model ids websocket/bin/funasr-wss-server-2pass.cpp:45-58),
run by onnxruntime 1.14.0 (websocket/onnxruntime-linux-x64-1.14.0/VERSION_NUMBER).  This file restates
that published architecture (SANM encoder, CifPredictorV2, ParaformerSANMDecoder; for config 3 the
CifPredictorV3 timestamp head, ContextualParaformerDecoder and the hotword Embedding+LSTM), batch 1, no padding,
which is what ORT executes for the reference's B=1 call.

PARITY UNPINNED for this file: the reference holds no golden vectors, tests or runnable model for the
encoder/predictor/decoder and onnxruntime/onnx are absent here, so the restatement is
anchored only on the reference's own call site contract (inputs speech[1,T,560] f32 + speech_lengths,
outputs logits[1,L,8404] f32 + token_num: paraformer.cpp:496-562) and on in-repo corroboration:
  * sqrt(512) input scale              onnxruntime/src/paraformer.h:28, paraformer-online.cpp:108
  * sinusoidal position encoding       onnxruntime/src/paraformer-online.cpp:240-268
  * CIF integrate-and-fire recurrence  onnxruntime/src/paraformer-online.cpp:270-345
  * cif_threshold 1.0, tail 0.45, 512-d, kernel 11 (fsmn_lorder 10)   onnxruntime/src/paraformer.h:112-123

Two pieces of this file ARE pinned against the reference's own compiled code (oracle/_ref/libfunasr_text_ref.so, built
from onnxruntime/src/paraformer-online.cpp where it lies; golden vectors tests/golden/cif_posenc_golden.npz): `cif`
(ParaformerOnline::CifSearch run in its offline form: token counts exact, frames to 1e-6) and `pos_enc`
(ParaformerOnline::GetPosEmb: 1e-6, 1e-5 at position 1000).  The INTERFACE of `forward` / `hotword_embed` (tensor layouts, token_num
semantics, the 4-output timestamp form, the [10, N, D] hotword output and its row selection) is pinned too: the reference's own
compiled Paraformer::Forward / CompileHotwordEmbedding run with this file as the network behind their sessions
(oracle/am_ref.py, oracle/fake_ort.cc) and produce the strings the oracle's host restatement predicts (tests/test_am_ref_cpu.py).

`emulate_bf16=True` (or "bf16" / "fp16") rounds to that 16-bit format exactly where the CUDA path stores
16-bit values (GEMM operands and the 16-bit activation buffers), keeping every reduction in fp32.  It is a
second oracle used to separate "kernel is wrong" from "operand rounding moved a value"; the acceptance
tolerances are stated against the plain fp32 oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict

import torch
import torch.nn.functional as F


@dataclass
class PfConfig:
    feat_dim: int = 560        # 80 mel x LFR 7   (paraformer.h:116, com-define PARA_LFR_M)
    d_model: int = 512         # encoder_size      (paraformer.h:117)
    n_heads: int = 4
    d_ff: int = 2048
    n_enc: int = 50            # encoders0[0] + encoders[0..48]
    n_dec: int = 16            # fsmn_layers       (paraformer.h:118)
    kernel: int = 11           # fsmn_lorder 10 +1 (paraformer.h:119)
    vocab: int = 8404
    cif_threshold: float = 1.0     # paraformer.h:121
    tail_threshold: float = 0.45   # paraformer.h:122
    pred_residual: int = 0     # 1: relu(conv(enc)+enc) (CifPredictor v1 / SURVEY a8); 0: relu(conv(enc)) (upstream CifPredictorV2)
    ln_eps: float = 1e-12
    # config 3 (SURVEY.md §8(a) a15/a16, Appendix B "Config-3 extras"); both default off = the 2-output model
    timestamp: int = 0         # 1: CifPredictorV3 upsample head -> us_alphas / us_cif_peak (4-output model, paraformer.cpp:549-563)
    contextual: int = 0        # 1: ContextualParaformerDecoder (bias_embed input, paraformer.cpp:515-531) + hotword LSTM (model_eb)
    us_times: int = 3          # upsample_times; TIME_RATE in util.cpp:851 encodes the x3
    smooth_factor2: float = 0.25
    noise_threshold2: float = 0.01

    def to_dict(self):
        return asdict(self)

    @classmethod
    def from_dict(cls, d):
        """Build from a model-file config dict (values arrive as floats); unknown keys are ignored."""
        f = cls.__dataclass_fields__
        return cls(**{k: (float(v) if isinstance(f[k].default, float) else int(v)) for k, v in d.items() if k in f})


def param_shapes(cfg: PfConfig):
    """name -> shape, in upstream FunASR state_dict naming."""
    D, Fd, V, K = cfg.d_model, cfg.d_ff, cfg.vocab, cfg.kernel
    out = {}

    def ln(p, n):
        out[p + ".weight"] = (n,)
        out[p + ".bias"] = (n,)

    def lin(p, o, i, bias=True):
        out[p + ".weight"] = (o, i)
        if bias:
            out[p + ".bias"] = (o,)

    for l in range(cfg.n_enc):
        p = "encoder.encoders0.0" if l == 0 else f"encoder.encoders.{l - 1}"
        din = cfg.feat_dim if l == 0 else D
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
    if cfg.timestamp:
        out["predictor.upsample_cnn.weight"] = (D, D, cfg.us_times)     # ConvTranspose1d: [in, out, k], stride k
        out["predictor.upsample_cnn.bias"] = (D,)
        for sfx in ("", "_reverse"):
            out["predictor.blstm.weight_ih_l0" + sfx] = (4 * D, D)
            out["predictor.blstm.weight_hh_l0" + sfx] = (4 * D, D)
            out["predictor.blstm.bias_ih_l0" + sfx] = (4 * D,)
            out["predictor.blstm.bias_hh_l0" + sfx] = (4 * D,)
        lin("predictor.cif_output2", 1, 2 * D)
    dec_names = [f"decoder.decoders.{l}" for l in range(cfg.n_dec - (1 if cfg.contextual else 0))]
    if cfg.contextual:
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


def _rb(x, on):
    """Round to the CUDA path's 16-bit operand format where `on` says so: False / None = fp32 (no rounding), True / "bf16" =
    bfloat16, "fp16" = IEEE half saturating at +-65504 like the device's cvt.rn.satfinite."""
    if not on:
        return x
    if on is True or on == "bf16":
        return x.bfloat16().float()
    if on == "fp16":
        return x.clamp(-65504.0, 65504.0).half().float()
    raise ValueError("unknown emulation mode %r" % (on,))


def pos_enc(T: int, depth: int) -> torch.Tensor:
    """SinusoidalPositionEncoder; same arithmetic as ParaformerOnline::GetPosEmb
    (paraformer-online.cpp:240-268): positions from 1, timescale step ln(1e4)/(depth/2-1)."""
    half = depth // 2
    inc = math.log(10000.0) / (half - 1)
    inv = torch.exp(torch.arange(half, dtype=torch.float32) * (-inc))
    pos = torch.arange(1, T + 1, dtype=torch.float32)
    st = pos[:, None] * inv[None, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


def _ln(x, W, p, eps):
    return F.layer_norm(x, (x.shape[-1],), W[p + ".weight"], W[p + ".bias"], eps)


def _lin(x, W, p, emu, bias=True):
    q = W.get(p + ".__int8__")
    if q is not None:          # int8 stand-in (quantize_dynamic_int8): the module carries its own bias
        return q(x)
    y = _rb(x, emu) @ _rb(W[p + ".weight"], emu).t()
    if bias and (p + ".bias") in W:
        y = y + W[p + ".bias"]
    return y


def quantize_dynamic_int8(W):
    """STAND-IN for the reference's deployed `model_quant.onnx` (offline-stream.cpp:75-77; websocket/run_server_offline.sh
    deploys quantize=true): onnxruntime's quantize_dynamic turns the MatMul / Gemm nodes into dynamic int8 kernels (per-tensor
    uint8 activations, int8 weights).  Neither onnxruntime nor the exported graph exist here, so the same matmuls of this
    restatement -- every nn.Linear-shaped projection that goes through `_lin` -- are replaced by PyTorch's dynamic int8 Linear
    (torch.ao.nn.quantized.dynamic.Linear, fbgemm / oneDNN kernels).  It measures PyTorch's int8 kernels, not ORT's: a labelled
    stand-in for the CPU baseline's int8 row, never a parity oracle.  Returns a new weight dict."""
    import warnings
    import torch.nn as nn
    out = dict(W)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from torch.ao.quantization import quantize_dynamic
        for name, w in W.items():
            if not name.endswith(".weight") or w.dim() != 2 or name in ("predictor.cif_output.weight", "predictor.cif_output2.weight", "bias_embed.weight"):
                continue
            if "weight_ih" in name or "weight_hh" in name:
                continue
            pfx = name[:-7]
            lin = nn.Linear(w.shape[1], w.shape[0], bias=(pfx + ".bias") in W)
            with torch.no_grad():
                lin.weight.copy_(w)
                if lin.bias is not None:
                    lin.bias.copy_(W[pfx + ".bias"])
            out[pfx + ".__int8__"] = quantize_dynamic(nn.Sequential(lin), {nn.Linear}, dtype=torch.qint8)[0]
    return out


def _fsmn(v, w, K):
    """depthwise conv1d, kernel K, zero pad (K-1)//2 left / K-1-left right, no bias, + identity."""
    left = (K - 1) // 2
    x = v.t()[None]                      # [1, D, T]
    x = F.pad(x, (left, K - 1 - left))
    y = F.conv1d(x, w, None, groups=w.shape[0])[0].t()
    return y + v


def encoder(feats: torch.Tensor, W, cfg: PfConfig, emu=False, taps=None):
    """feats [T,560] (LFR+CMVN output) -> enc [T,512].  SURVEY.md §8(a) a7 / Appendix B."""
    D, H = cfg.d_model, cfg.n_heads
    T = feats.shape[0]
    x = feats * (D ** 0.5) + pos_enc(T, cfg.feat_dim)
    if taps is not None:
        taps["enc_in"] = x.clone()
    scale = (D // H) ** -0.5
    for l in range(cfg.n_enc):
        p = "encoder.encoders0.0" if l == 0 else f"encoder.encoders.{l - 1}"
        h = _ln(x, W, p + ".norm1", cfg.ln_eps)
        qkv = _rb(_lin(h, W, p + ".self_attn.linear_q_k_v", emu), emu)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        mem = _rb(_fsmn(v, W[p + ".self_attn.fsmn_block.weight"], cfg.kernel), emu)
        # 128^-1/2 is applied to the fp32 scores (upstream scales q first; same value to rounding)
        att = _mha_scaled(q, k, v, H, scale, emu)
        att = _rb(att, emu)
        y = _lin(att, W, p + ".self_attn.linear_out", emu) + mem
        x = x + y if l > 0 else y
        h2 = _ln(x, W, p + ".norm2", cfg.ln_eps)
        f1 = _rb(torch.relu(_lin(h2, W, p + ".feed_forward.w_1", emu)), emu)
        x = x + _lin(f1, W, p + ".feed_forward.w_2", emu)
        if taps is not None and (l == 0 or l == cfg.n_enc - 1):
            taps[f"enc_l{l}"] = x.clone()
    enc = _ln(x, W, "encoder.after_norm", cfg.ln_eps)
    return enc


def _mha_scaled(q, k, v, H, scale, emu):
    Tq, D = q.shape
    dk = D // H
    qh = q.reshape(Tq, H, dk).transpose(0, 1)
    kh = k.reshape(-1, H, dk).transpose(0, 1)
    vh = v.reshape(-1, H, dk).transpose(0, 1)
    s = (qh @ kh.transpose(1, 2)) * scale
    p = torch.softmax(s, dim=-1)
    o = _rb(p, emu) @ vh
    return o.transpose(0, 1).reshape(Tq, D)


def predictor_alphas(enc, W, cfg: PfConfig, emu=False):
    """CifPredictorV2 alphas.  enc [T,512] -> alpha [T]."""
    x = _rb(enc, emu).t()[None]
    x = F.pad(x, (1, 1))
    w = _rb(W["predictor.cif_conv1d.weight"], emu)
    c = F.conv1d(x, w, W["predictor.cif_conv1d.bias"])[0].t()
    if cfg.pred_residual:
        c = c + enc
    o = torch.relu(c)
    logit = o @ W["predictor.cif_output.weight"].t() + W["predictor.cif_output.bias"]
    alpha = torch.sigmoid(logit)[:, 0]
    alpha = torch.relu(alpha * 1.0 - 0.0)       # smooth_factor 1.0, noise_threshold 0
    return alpha


def cif(hidden, alphas, threshold: float):
    """Integrate-and-fire, literal sequential recurrence (same as ParaformerOnline::CifSearch,
    paraformer-online.cpp:270-345, and upstream `cif`): returns embeds [L,D], fires [T'] (the value of
    `integrate` after adding alpha_t, before the subtraction)."""
    Tn, D = hidden.shape
    integrate = torch.zeros((), dtype=torch.float32)
    frame = torch.zeros(D, dtype=torch.float32)
    one = torch.ones((), dtype=torch.float32)
    fires = torch.zeros(Tn, dtype=torch.float32)
    out = []
    for t in range(Tn):
        a = alphas[t]
        dc = one - integrate
        integrate = integrate + a
        fires[t] = integrate
        fire = bool(integrate >= threshold)
        cur = dc if fire else a
        remainds = a - cur
        frame = frame + cur * hidden[t]
        if fire:
            out.append(frame.clone())
            integrate = integrate - one
            frame = remainds * hidden[t]
    emb = torch.stack(out, 0) if out else torch.zeros(0, D)
    return emb, fires


def predictor(enc, W, cfg: PfConfig, emu=False):
    """alphas -> tail -> CIF.  Returns embeds [L,512], token_num (int), alphas' [T+1], fires [T+1]."""
    alpha = predictor_alphas(enc, W, cfg, emu)
    # tail_process_fn (batch-1, mask all ones): append one frame with alpha = tail_threshold, hidden 0
    alpha_t = torch.cat([alpha, torch.tensor([cfg.tail_threshold], dtype=torch.float32)])
    hidden = torch.cat([_rb(enc, emu), torch.zeros(1, enc.shape[1])], 0)
    token_num = int(torch.floor(alpha_t.sum()).item())
    emb, fires = cif(hidden, alpha_t, cfg.cif_threshold)
    return emb, token_num, alpha_t, fires


def lstm(x, W, prefix, sfx="", reverse=False, emu=False):
    """One direction of a single-layer torch.nn.LSTM (gate order i, f, g, o), literal recurrence.
    x [T, I] -> h [T, H].  With emu the input projection, the hidden state fed back into the recurrence and the
    output are rounded to bf16 where the CUDA path stores bf16; the cell state stays fp32."""
    w_ih, w_hh = W[prefix + ".weight_ih_l0" + sfx], W[prefix + ".weight_hh_l0" + sfx]
    b = W[prefix + ".bias_ih_l0" + sfx] + W[prefix + ".bias_hh_l0" + sfx]
    H = w_hh.shape[1]
    gx = _rb(_rb(x, emu) @ _rb(w_ih, emu).t() + b, emu)
    whh_t = _rb(w_hh, emu).t()
    h = torch.zeros(H)
    c = torch.zeros(H)
    out = torch.zeros(x.shape[0], H)
    order = range(x.shape[0] - 1, -1, -1) if reverse else range(x.shape[0])
    for t in order:
        g = gx[t] + _rb(h, emu) @ whh_t
        i, f, gg, o = g[:H], g[H:2 * H], g[2 * H:3 * H], g[3 * H:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return _rb(out, emu)


def cif_wo_hidden(alphas, threshold):
    """Upstream `cif_wo_hidden`: running sum, value recorded BEFORE the subtraction, subtract `threshold` on fire.
    The consumer is TimestampOnnx (util.cpp:866-870: a peak is `> 1 - 1e-4`)."""
    th = torch.tensor(threshold, dtype=torch.float32)
    integrate = torch.zeros((), dtype=torch.float32)
    fires = torch.zeros(alphas.shape[0], dtype=torch.float32)
    for t in range(alphas.shape[0]):
        integrate = integrate + alphas[t]
        fires[t] = integrate
        if bool(integrate >= th):
            integrate = integrate - th
    return fires


def upsample_timestamp(enc, token_num, W, cfg: PfConfig, emu=False):
    """CifPredictorV3.get_upsample_timestmap, upsample_type cnn_blstm, use_cif1_cnn false (SURVEY.md Appendix B
    "Config-3 extras"; consumer contract util.cpp:838-870, paraformer.cpp:549-563).  enc [T,512] ->
    us_alphas [3T], us_cif_peak [3T]."""
    T, D = enc.shape
    k = cfg.us_times
    wt = _rb(W["predictor.upsample_cnn.weight"], emu)            # ConvTranspose1d weight [in, out, k], stride k
    x = _rb(enc, emu)
    up = torch.stack([x @ wt[:, :, j] for j in range(k)], 1) + W["predictor.upsample_cnn.bias"]   # [T, k, D]
    up = _rb(up.reshape(T * k, D), emu)
    hcat = torch.cat([lstm(up, W, "predictor.blstm", "", False, emu), lstm(up, W, "predictor.blstm", "_reverse", True, emu)], 1)
    a2 = torch.sigmoid(hcat @ W["predictor.cif_output2.weight"].t() + W["predictor.cif_output2.bias"])[:, 0]
    a2 = torch.relu(a2 * cfg.smooth_factor2 - cfg.noise_threshold2)
    tok = a2.sum()
    a2 = a2 * (torch.tensor(float(token_num), dtype=torch.float32) / tok)
    peaks = cif_wo_hidden(a2, cfg.cif_threshold - 1e-4)
    return a2, peaks


def hotword_embed(ids, W, emu=False):
    """model_eb.onnx (paraformer.cpp:592-693): hotword ids [N, 10] -> Embedding -> LSTM -> [10, N, 512]."""
    ids = torch.as_tensor(ids, dtype=torch.long)
    emb = W["bias_embed.weight"][ids]                                  # [N, 10, D]
    out = torch.stack([lstm(emb[n], W, "bias_encoder", "", False, emu) for n in range(emb.shape[0])], 1)
    return out                                                         # [10, N, D]


def select_hotword_rows(hw_out, lengths):
    """CompileHotwordEmbedding's pick of step len_j - 1 for word j (paraformer.cpp:676-682)."""
    return torch.stack([hw_out[int(lengths[j]) - 1, j] for j in range(hw_out.shape[1])], 0)


def decoder(emb, enc, token_num, W, cfg: PfConfig, emu=False, taps=None, hw_emb=None):
    """ParaformerSANMDecoder; with cfg.contextual the ContextualParaformerDecoder
    (a16: the last attention layer is `last_decoder`, its self-attention output attends over the hotword
    embeddings through `bias_decoder`, and `bias_output` (1x1 conv over [x_src_attn ; cx]) replaces the
    cross-attention residual).  emb [L,512], enc [T,512], hw_emb [N,512] -> logits [L,V] (before log_softmax)."""
    D, H = cfg.d_model, cfg.n_heads
    L = emb.shape[0]
    scale = (D // H) ** -0.5
    tmask = (torch.arange(L) < token_num).float()[:, None]
    y = emb
    memory = _rb(enc, emu)

    def ffn(t, p):
        f1 = torch.relu(_lin(t, W, p + ".feed_forward.w_1", emu))
        f1 = _ln(_rb(f1, emu), W, p + ".feed_forward.norm", cfg.ln_eps)
        return _lin(f1, W, p + ".feed_forward.w_2", emu, bias=False)

    def self_block(y, p):
        t = ffn(_ln(y, W, p + ".norm1", cfg.ln_eps), p)
        t2 = _rb(_ln(t, W, p + ".norm2", cfg.ln_eps), emu) * tmask
        m = _fsmn(t2, W[p + ".self_attn.fsmn_block.weight"], cfg.kernel) * tmask
        return y + m

    def cross(y, p_norm, p_att, mem):
        q = _rb(_lin(_ln(y, W, p_norm, cfg.ln_eps), W, p_att + ".linear_q", emu), emu)
        kv = _rb(_lin(mem, W, p_att + ".linear_k_v", emu), emu)
        att = _rb(_mha_scaled(q, kv[:, :D], kv[:, D:], H, scale, emu), emu)
        return _lin(att, W, p_att + ".linear_out", emu)

    n_plain = cfg.n_dec - (1 if cfg.contextual else 0)
    for l in range(n_plain):
        p = f"decoder.decoders.{l}"
        y = self_block(y, p)
        y = y + cross(y, p + ".norm3", p + ".src_attn", memory)
        if taps is not None and l == 0:
            taps["dec_l0"] = y.clone()
    if cfg.contextual:
        p = "decoder.last_decoder"
        x_self = self_block(y, p)
        x_src = cross(x_self, p + ".norm3", p + ".src_attn", memory)
        cx = cross(x_self, "decoder.bias_decoder.norm3", "decoder.bias_decoder.src_attn", _rb(torch.as_tensor(hw_emb, dtype=torch.float32), emu))
        cat = _rb(torch.cat([x_src, cx], 1), emu)
        y = x_self + cat @ _rb(W["decoder.bias_output.weight"][:, :, 0], emu).t()
        if taps is not None:
            taps["dec_ctx"] = y.clone()
    p = "decoder.decoders3.0"
    y = ffn(_ln(y, W, p + ".norm1", cfg.ln_eps), p)
    y = _ln(y, W, "decoder.after_norm", cfg.ln_eps)
    if taps is not None:
        taps["dec_out"] = y.clone()
    logits = _lin(y, W, "decoder.output_layer", emu)
    return logits


@torch.no_grad()
def forward(feats, W, cfg: PfConfig, emulate_bf16=False, want_taps=True, hw_emb=None):
    """feats: float32 [T,560] (numpy or tensor).  Returns dict of taps:
    enc [T,512], alphas [T+1], fires [T+1], embeds [L,512], token_num, logits [L,V] (raw),
    logprobs [L,V] (= what the reference graph outputs), ids (FindMax over the first token_num rows);
    with cfg.timestamp also us_alphas [3T], us_peaks [3T]; cfg.contextual needs hw_emb [N,512]."""
    feats = torch.as_tensor(feats, dtype=torch.float32)
    taps = {} if want_taps else None
    enc = encoder(feats, W, cfg, emulate_bf16, taps)
    emb, token_num, alphas, fires = predictor(enc, W, cfg, emulate_bf16)
    out = dict(taps or {})
    out.update(enc=enc, alphas=alphas, fires=fires, embeds=emb, token_num=token_num)
    if cfg.timestamp:
        out["us_alphas"], out["us_peaks"] = upsample_timestamp(enc, token_num, W, cfg, emulate_bf16)
    if emb.shape[0] == 0:
        out.update(logits=torch.zeros(0, cfg.vocab), logprobs=torch.zeros(0, cfg.vocab), ids=[])
        return out
    logits = decoder(emb, enc, token_num, W, cfg, emulate_bf16, taps, hw_emb)
    if taps:
        out.update(taps)
    out["logits"] = logits
    out["logprobs"] = torch.log_softmax(logits, dim=-1)
    n = min(token_num, logits.shape[0])
    # FindMax: first maximum wins (util.cpp:63-74) == torch.argmax on CPU? not guaranteed -> do it literally
    ids = []
    lp = out["logprobs"]
    for i in range(n):
        row = lp[i]
        m = row.max()
        ids.append(int((row == m).nonzero()[0, 0]))
    out["ids"] = ids
    return out


def logprob_topk(logits, k: int):
    """What the reference's log-prob consumers read from the graph's log_softmax output (WfstDecoder::Search,
    wfst-decoder.cpp:27-57), pruned to the k best entries per row.  Order: value descending, index ascending, so
    entry 0 is FindMax's first-max-wins argmax (util.cpp:63-74).  Returns (lse [L], logprob [L,k], ids [L,k])."""
    import numpy as np
    x = torch.as_tensor(logits, dtype=torch.float32)
    lse = torch.logsumexp(x, dim=-1)
    lp = (x - lse[:, None]).numpy()
    xn = x.numpy()
    ids = np.zeros((x.shape[0], k), np.int64)
    for r in range(x.shape[0]):
        order = np.lexsort((np.arange(x.shape[1]), -xn[r]))   # primary: -value, secondary: index
        ids[r] = order[:k]
    return lse.numpy(), np.take_along_axis(lp, ids, 1), ids


def flops(T: int, L: int, cfg: PfConfig = PfConfig()) -> float:
    """Algorithmic FLOPs of one segment: 2*M*N*K per contraction."""
    D, Fd, V = cfg.d_model, cfg.d_ff, cfg.vocab
    enc = 0.0
    for l in range(cfg.n_enc):
        din = cfg.feat_dim if l == 0 else D
        enc += 2 * T * din * 3 * D + 2 * T * D * D + 4 * T * D * Fd + 4 * T * T * D
    pred = 2 * T * D * D * 3 + 2 * T * D
    dec = cfg.n_dec * (4 * L * D * Fd + 4 * L * D * D + 4 * T * D * D + 4 * L * T * D) + 4 * L * D * Fd + 2 * L * D * V
    return enc + pred + dec
