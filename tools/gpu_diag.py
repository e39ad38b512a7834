"""GPU bring-up diagnostics: every kernel against numpy / the oracle, one group per process so a CUDA fault
in one group cannot poison the others.   python tools/gpu_diag.py [group ...]   (no args = all groups)
"""
import importlib
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["gemm", "conv3", "layernorm", "attn_check", "attn", "fsmn", "cif", "frontend", "e2e_small", "e2e_full"]


def bf(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).bfloat16().float().numpy()


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def g_gemm(capi):
    rng = np.random.default_rng(0)
    for (M, N, K) in [(128, 256, 64), (128, 256, 512), (300, 512, 560), (1000, 1536, 512), (4096, 2048, 512), (513, 512, 2048)]:
        A = rng.standard_normal((M, K)).astype(np.float32)
        W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
        bias = rng.standard_normal(N).astype(np.float32)
        ref = bf(A) @ bf(W).T
        out = capi.op_gemm(A, W)
        print("gemm plain", (M, N, K), "rel err", rel(out, ref), flush=True)
        if rel(out, ref) > 1e-3:
            bad = np.abs(out - ref) > 1e-2 * np.abs(ref).max()
            rows = np.where(bad.any(1))[0]
            cols = np.where(bad.any(0))[0]
            print("   bad rows", rows[:10], "n", len(rows), "bad cols", cols[:10], "n", len(cols))
            print("   out[0,:8]", out[0, :8], "ref", ref[0, :8])
        add = rng.standard_normal((M, N)).astype(np.float32)
        res = rng.standard_normal((M, N)).astype(np.float32)
        out = capi.op_gemm(A, W, bias=bias, add=add, res=res, relu=1)
        ref2 = np.maximum(ref + bias, 0) + bf(add) + res
        print("gemm bias+relu+add+res", rel(out, ref2))
        out = capi.op_gemm(A, W, bias=bias, res=res, relu=2, out_bf16=True)
        ref3 = bf(np.maximum(ref + bias + res, 0))
        print("gemm relu2 bf16-out", rel(out, ref3))
    M, N, K = 257, 8404, 512
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32) * 0.1
    out, am = capi.op_gemm(A, W, bias=bias, argmax=True)
    ref = bf(A) @ bf(W).T + bias
    print("gemm vocab", rel(out, ref), "argmax vs own logits", int((am != out.argmax(1)).sum()), "vs ref", int((am != ref.argmax(1)).sum()))


def g_conv3(capi):
    import torch
    rng = np.random.default_rng(1)
    # 3 segments with one zero gap row after each
    lens = [50, 130, 7]
    M = sum(lens) + len(lens)
    X = np.zeros((M, 512), np.float32)
    r = 0
    segs = []
    for L in lens:
        X[r:r + L] = rng.standard_normal((L, 512))
        segs.append((r, L))
        r += L + 1
    w = (rng.standard_normal((512, 512, 3)) / np.sqrt(1536)).astype(np.float32)
    b = rng.standard_normal(512).astype(np.float32)
    Wr = np.ascontiguousarray(w.transpose(0, 2, 1).reshape(512, 1536))
    out = capi.op_conv3(X, Wr, b)
    worst = 0
    for (r0, L) in segs:
        xs = torch.from_numpy(bf(X[r0:r0 + L])).t()[None]
        ref = torch.nn.functional.conv1d(torch.nn.functional.pad(xs, (1, 1)), torch.from_numpy(bf(w)), torch.from_numpy(b))[0].t().numpy()
        worst = max(worst, rel(out[r0:r0 + L], ref))
    print("conv3 rel err", worst)


def g_layernorm(capi):
    import torch
    rng = np.random.default_rng(2)
    for D in (512, 560, 2048):
        x = (rng.standard_normal((77, D)) * 3 + 1).astype(np.float32)
        g = rng.standard_normal(D).astype(np.float32)
        b = rng.standard_normal(D).astype(np.float32)
        ref = torch.nn.functional.layer_norm(torch.from_numpy(x), (D,), torch.from_numpy(g), torch.from_numpy(b), 1e-12).numpy()
        o32, o16 = capi.op_layernorm(x, g, b)
        print("layernorm f32-in", D, "abs err", float(np.abs(o32 - ref).max()), "bf16 out err", float(np.abs(o16 - bf(ref)).max()))
        xb = bf(x)
        refb = torch.nn.functional.layer_norm(torch.from_numpy(xb), (D,), torch.from_numpy(g), torch.from_numpy(b), 1e-12).numpy()
        o32, o16 = capi.op_layernorm(x, g, b, in_bf16=True)
        print("layernorm bf16-in", D, "abs err", float(np.abs(o32 - refb).max()))


def _attn_case(rng, q_lens, kv_lens, H=4):
    D = H * 128
    q_off = np.concatenate([[0], np.cumsum(q_lens)[:-1]]).astype(np.int32)
    kv_off = np.concatenate([[0], np.cumsum(np.asarray(kv_lens) + 1)[:-1]]).astype(np.int32)  # +1 gap row like the engine
    q = rng.standard_normal((int(sum(q_lens)), D)).astype(np.float32)
    k = rng.standard_normal((int(sum(kv_lens) + len(kv_lens)), D)).astype(np.float32)
    v = rng.standard_normal((int(sum(kv_lens) + len(kv_lens)), D)).astype(np.float32)
    return q, k, v, q_off, kv_off


def _attn_ref(q, k, v, q_off, q_len, kv_off, kv_len, H=4):
    out = np.zeros_like(q)
    qb, kb, vb = bf(q), bf(k), bf(v)
    for s in range(len(q_len)):
        for h in range(H):
            qs = qb[q_off[s]:q_off[s] + q_len[s], h * 128:(h + 1) * 128]
            ks = kb[kv_off[s]:kv_off[s] + kv_len[s], h * 128:(h + 1) * 128]
            vs = vb[kv_off[s]:kv_off[s] + kv_len[s], h * 128:(h + 1) * 128]
            sc = (qs @ ks.T) * (128 ** -0.5)
            sc = sc - sc.max(1, keepdims=True)
            p = np.exp(sc)
            p /= p.sum(1, keepdims=True)
            out[q_off[s]:q_off[s] + q_len[s], h * 128:(h + 1) * 128] = p @ vs
    return out


def g_attn_impl(capi, impl):
    rng = np.random.default_rng(3)
    for q_lens, kv_lens in [([33], [33]), ([64], [64]), ([128], [128]), ([167], [167]), ([200, 1, 64, 129], [200, 1, 64, 129]),
                            ([40, 90], [83, 167]), ([1000], [1000])]:
        q, k, v, q_off, kv_off = _attn_case(rng, q_lens, kv_lens)
        ref = _attn_ref(q, k, v, q_off, q_lens, kv_off, kv_lens)
        out = capi.op_attention(q, k, v, q_off, q_lens, kv_off, kv_lens, impl=impl)
        e = rel(out, ref)
        print("attention impl", impl, q_lens, kv_lens, "rel err", e, flush=True)
        if e > 2e-2:
            d = np.abs(out - ref)
            rows = np.where((d > 0.05 * np.abs(ref).max()).any(1))[0]
            print("   bad rows", rows[:16], "n", len(rows), " out[0,:4]", out[0, :4], "ref", ref[0, :4])


def g_fsmn(capi):
    import torch
    rng = np.random.default_rng(4)
    lens = [1, 5, 11, 40, 300]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    x = rng.standard_normal((off[-1], 512)).astype(np.float32)
    w = (rng.standard_normal((512, 1, 11)) * 0.3).astype(np.float32)
    out = capi.op_fsmn(x, w, off)
    worst = 0
    for s, L in enumerate(lens):
        xs = torch.from_numpy(bf(x[off[s]:off[s + 1]]))
        y = torch.nn.functional.conv1d(torch.nn.functional.pad(xs.t()[None], (5, 5)), torch.from_numpy(w), groups=512)[0].t() + xs
        worst = max(worst, float(np.abs(out[off[s]:off[s + 1]] - bf(y.numpy())).max()))
    print("fsmn abs err (vs bf16-rounded ref)", worst)


def g_cif(capi):
    import torch
    from oracle import paraformer_ref as R
    rng = np.random.default_rng(5)
    lens = [3, 40, 167, 1000]
    off = np.concatenate([[0], np.cumsum(np.asarray(lens) + 1)]).astype(np.int32)
    alphas = rng.uniform(0.05, 0.95, off[-1]).astype(np.float32)
    hidden = rng.standard_normal((off[-1], 512)).astype(np.float32)
    for s in range(len(lens)):
        alphas[off[s + 1] - 1] = 0.45
        hidden[off[s + 1] - 1] = 0
    n_tok, fires, emb, ff = capi.op_cif(alphas, hidden, off)
    t0 = 0
    for s in range(len(lens)):
        e_ref, f_ref = R.cif(torch.from_numpy(hidden[off[s]:off[s + 1]]), torch.from_numpy(alphas[off[s]:off[s + 1]]), 1.0)
        L = e_ref.shape[0]
        fr = np.where(f_ref.numpy() >= 1.0)[0]
        print("cif seg", s, "tokens", n_tok[s], L, "fires bit-exact", bool((fires[off[s]:off[s + 1]] == f_ref.numpy()).all()),
              "embeds bit-exact", bool(n_tok[s] == L and (emb[t0:t0 + L] == e_ref.numpy()).all()),
              "frames ok", bool(n_tok[s] == L and (ff[t0:t0 + L] == fr).all()))
        t0 += n_tok[s]


def _model(tmp, cfg_over, jitter=True):
    synth = importlib.import_module("asr-2pass_b200.synth")
    cfg, W, means, vars_, toks = synth.write_synthetic_model_dir(tmp, cfg_over, seed=0, jitter_ln=jitter)
    return synth, cfg, W, means, vars_, toks


def g_frontend(capi):
    import tempfile
    from oracle import frontend as F
    tmp = tempfile.mkdtemp()
    synth, cfg, W, means, vars_, toks = _model(tmp, dict(n_enc=1, n_dec=1))
    eng = capi.Engine(tmp, max_rows=4096, max_segments=64)
    for n in (400, 559, 560, 16000, 160000, 960000):
        pcm = synth.make_audio(n, 77 + n)
        fb, feats = eng.frontend(pcm)
        pf = pcm.astype(np.float32) / np.float32(32768)
        fb_o = F.fbank(pf)
        ft_o = F.lfr_cmvn(fb_o, means, vars_)
        msg = "frontend n=%d fbank max abs err %.3g (bit-equal %.4f)  feats err %.3g" % (
            n, np.abs(fb - fb_o).max(), (fb == fb_o).mean(), np.abs(feats - ft_o).max())
        if F.ref_lib() is not None:
            msg += "  | vs reference knf: %.3g" % np.abs(fb - F.fbank_ref(pf)).max()
        print(msg, flush=True)


def _e2e(capi, cfg_over, seg_secs, max_rows):
    import tempfile
    import torch
    from oracle import frontend as F
    from oracle import paraformer_ref as R
    tmp = tempfile.mkdtemp()
    synth, cfg, W, means, vars_, toks = _model(tmp, cfg_over)
    pc = R.PfConfig.from_dict(cfg)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    eng = capi.Engine(tmp, max_rows=max_rows, max_segments=64)
    eng.set_option("taps", 1)
    lens = [int(s * 16000) for s in seg_secs]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pcm = np.concatenate([synth.make_audio(n, 100 + i) for i, n in enumerate(lens)])
    b = capi.Batch(eng, int(offs[-1]) + 16)
    t = time.time()
    res = b.forward_s16(pcm, offs)
    print("forward ok in %.3fs; launches %d; tokens %s; T %s" % (time.time() - t, b.launches, res["token_counts"], res["lfr_frames"]), flush=True)
    for i in range(len(lens)):
        pf = pcm[offs[i]:offs[i + 1]].astype(np.float32) / np.float32(32768)
        feats = F.lfr_cmvn(F.fbank(pf), means, vars_)
        if feats.shape[0] == 0:
            print("seg %d: too short, tokens %d" % (i, res["token_counts"][i]))
            continue
        o = R.forward(feats, Wt, pc)
        o2 = R.forward(feats, Wt, pc, emulate_bf16=True)
        msg = ["seg %d T=%d" % (i, feats.shape[0])]
        for name in ("feats", "enc", "alphas", "embeds", "logits"):
            try:
                g = b.tap(name, i)
            except Exception as ex:  # noqa
                msg.append("%s: tap failed %s" % (name, ex))
                continue
            if name == "feats":
                msg.append("feats abs %.2g" % np.abs(g - feats).max())
                continue
            a, a2 = o[name].numpy(), o2[name].numpy()
            if g.shape != a.shape:
                msg.append("%s shape %s vs %s" % (name, g.shape, a.shape))
                n = min(len(g), len(a))
                g, a, a2 = g[:n], a[:n], a2[:n]
            msg.append("%s rel(fp32) %.3g rel(emu) %.3g" % (name, rel(g, a), rel(g, a2)))
        ids = res["token_ids"][res["token_offsets"][i]:res["token_offsets"][i + 1]]
        fr = res["fire_frames"][res["token_offsets"][i]:res["token_offsets"][i + 1]]
        fr_o = np.where(o["fires"].numpy() >= 1.0)[0]
        fr_o2 = np.where(o2["fires"].numpy() >= 1.0)[0]
        msg.append("tokens gpu %d fp32 %d emu %d" % (len(ids), len(o["ids"]), len(o2["ids"])))
        if len(fr) == len(fr_o):
            msg.append("fire mismatches fp32 %d emu %d" % (int((fr != fr_o).sum()), int((fr != fr_o2).sum()) if len(fr) == len(fr_o2) else -1))
        if len(ids) == len(o["ids"]):
            msg.append("id mismatches fp32 %d emu %d" % (sum(int(a != b_) for a, b_ in zip(ids, o["ids"])),
                                                         sum(int(a != b_) for a, b_ in zip(ids, o2["ids"])) if len(ids) == len(o2["ids"]) else -1))
        print("  ".join(msg), flush=True)


def g_e2e_small(capi):
    _e2e(capi, dict(n_enc=2, n_dec=2), [1.0, 3.3, 10.0, 0.02, 5.25], 2048)


def g_e2e_full(capi):
    _e2e(capi, dict(), [10.0, 2.0, 20.0], 4096)


def main():
    args = sys.argv[1:]
    if not args:
        for g in GROUPS:
            print("==== %s ====" % g, flush=True)
            r = subprocess.run(["timeout", "600", sys.executable, os.path.abspath(__file__), g])
            print("---- %s exit %d" % (g, r.returncode), flush=True)
        return
    capi = importlib.import_module("asr-2pass_b200.capi")
    print("devices:", capi.device_count())
    for g in args:
        fn = {"gemm": g_gemm, "conv3": g_conv3, "layernorm": g_layernorm, "attn_check": lambda c: g_attn_impl(c, 1),
              "attn": lambda c: g_attn_impl(c, 0), "fsmn": g_fsmn, "cif": g_cif, "frontend": g_frontend,
              "e2e_small": g_e2e_small, "e2e_full": g_e2e_full}[g]
        fn(capi)


if __name__ == "__main__":
    main()
