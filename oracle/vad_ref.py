"""ORACLE (test infrastructure only) - fp32 CPU restatement of the FSMN-VAD acoustic model the reference runs through
onnxruntime in FsmnVad::Forward (onnxruntime/src/fsmn-vad.cpp:72-135: inputs feats [1,T,400] + four caches [1,128,19,1],
outputs scores [1,T,248] + the updated caches), SURVEY.md §8(f) rank 2.

The graph is a third-party artefact that is NOT under /root/reference (FunASR export of
`speech_fsmn_vad_zh-cn-16k-common-pytorch`); this file restates the published architecture (FunASR
funasr/models/fsmn_vad_streaming/encoder.py `FSMN`): in_linear1 400->140, in_linear2 140->250, ReLU, 4 x {linear 250->128
(no bias), FSMN memory block (depthwise causal conv, lorder 20, stride 1, + identity), affine 128->250, ReLU}, out_linear1
250->140, out_linear2 140->248, softmax.  PARITY UNPINNED (no runnable reference graph, no golden vectors); anchored on the
reference's call-site contract only: feature dim 400 = 80 mel x LFR 5 (com-define.h:103-109), four caches of 128 x 19
(fsmn-vad.cpp:96-100,127-133), output dim consumed as scores[t][sil_pdf_id] with sil_pdf_ids = {0} (e2e-vad.h:602-608).

The front end (fbank 80, LFR m=5 n=1, CMVN) is FsmnVad::FbankKaldi / LfrCmvn (fsmn-vad.cpp:137-224): the fbank is the same
knf configuration as the acoustic model's (oracle/frontend.py), the LFR is restated below.  Front end, cache interface and the
equivalence "whole recording in one pass == the reference's 1 s pieces through FsmnVadOnline" ARE pinned: the reference's own
fsmn-vad.cpp / fsmn-vad-online.cpp, compiled in place and driven like Audio::CutSplit with `forward` behind the session, see the
same frames and features and cut the same segments (tests/test_vad.py::test_whole_recording_vad_equals_reference_cutsplit).
"""
import numpy as np
import torch

DIMS = dict(input_dim=400, input_affine_dim=140, linear_dim=250, proj_dim=128, lorder=20, n_layers=4, output_affine_dim=140, output_dim=248)


def param_shapes(d=DIMS):
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


def lfr_cmvn(fb, means, vars_, m=5, n=1):
    """FsmnVad::LfrCmvn (fsmn-vad.cpp:182-224): left-pad (m-1)/2 copies of frame 0, stack m frames every n, right-pad by
    repeating the last frame, then (x + mean) * var."""
    fb = np.asarray(fb, np.float32)
    T = fb.shape[0]
    if T == 0:
        return np.zeros((0, fb.shape[1] * m), np.float32)
    T_lfr = int(np.ceil(T / n))
    left = (m - 1) // 2
    idx = np.clip(np.arange(T_lfr)[:, None] * n + np.arange(m)[None, :] - left, 0, T - 1)
    x = fb[idx].reshape(T_lfr, -1)
    return ((x + np.asarray(means, np.float32)[None, :]).astype(np.float32) * np.asarray(vars_, np.float32)[None, :]).astype(np.float32)


def init_caches(d=DIMS):
    return [torch.zeros(d["proj_dim"], d["lorder"] - 1) for _ in range(d["n_layers"])]


@torch.no_grad()
def forward(feats, W, caches=None, d=DIMS):
    """feats [T,400] -> (scores [T,248] softmax, new caches).  With caches=None the left context is zero (start of audio).
    Chunked calls that pass the caches along equal one call over the concatenation (tested)."""
    x = torch.as_tensor(feats, dtype=torch.float32)
    W = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in W.items()}
    caches = init_caches(d) if caches is None else caches
    lin = lambda t, p, bias=True: t @ W[p + ".weight"].t() + (W[p + ".bias"] if bias else 0.0)
    x = lin(x, "encoder.in_linear1.linear")
    x = torch.relu(lin(x, "encoder.in_linear2.linear"))
    new_caches = []
    L = d["lorder"]
    for i in range(d["n_layers"]):
        p = "encoder.fsmn.%d" % i
        h = lin(x, p + ".linear.linear", bias=False)                      # [T,128]
        ycat = torch.cat([caches[i].t(), h], 0)                           # [19 + T, 128]
        new_caches.append(ycat[-(L - 1):].t().clone())
        w = W[p + ".fsmn_block.conv_left.weight"][:, 0, :, 0]             # [128, 20]
        T = h.shape[0]
        mem = torch.zeros_like(h)
        for k in range(L):
            mem = mem + ycat[k:k + T] * w[:, k][None, :]
        h = h + mem
        x = torch.relu(lin(h, p + ".affine.linear"))
    x = lin(x, "encoder.out_linear1.linear")
    x = lin(x, "encoder.out_linear2.linear")
    return torch.softmax(x, dim=-1), new_caches
