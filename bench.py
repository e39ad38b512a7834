#!/usr/bin/env python
"""bench.py — RTFx (audio seconds / wall second) of the offline Paraformer-large acoustic-model path
(`Paraformer::Forward` over VAD-cut segments) on N B200s, next to the CPU oracle timed on the host cores.

Workload = BASELINE.json configs[1]: 1024 synthetic VAD segments of 2-20 s (seeded speech-like noise),
length-sorted and packed into batches, random-init Paraformer-large weights (215.8 M parameters).
One "step" = one pass of the hot path over the whole workload (every rank runs the full per-GPU workload;
segments are independent, so there is no collective on the data path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--segments S]

Prints ONE JSON line.  Keys beyond the contract:
  value               PCM already resident in HBM (CUDA events)
  e2e                 through the C ABI with pinned HOST int16 buffers, H2D + D2H inside the timed region
  e2e_model_forward   through the reference-facing funasr::Model::Forward(float**, int*, ...) of the host library, pageable
                      float buffers in, text strings out (what a reference caller passes and receives)
  parity              the timed CUDA path checked against the fp32 oracle on the cpu_baseline sample (asr-2pass_b200/parity.py)
  cpu_baseline        the oracle port on the host cores, fixed sample, fp32 and the int8 stand-in, per-core figures
  config4 / config5   the PRODUCT's multi-GPU path (one handle, per-GPU queues: MultiGpuParaformer) driven by rank 0 over all N
                      GPUs on BASELINE.json configs[3] (1 h stream, arrival order) and configs[4] (256 x 60 s): strong scaling
  config3             BASELINE.json configs[2] on rank 0's GPU: contextual decoder + 100 hotwords + timestamp head on 256 segments
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "RTFx (audio s/s) Paraformer-large offline batched"
UNIT = "audio_s/s"
CPU_SAMPLE = 96          # segments of the fixed cpu_baseline / parity sample (the same indices on every box)
PARITY_TENSORS = 24      # of those, how many also carry encoder output and logits for the tensor comparisons


def rank_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


def lfr_frames(n):
    """T of a segment with n samples (feature-window.cc:73-87 snip_edges, paraformer.cpp:424) -- host arithmetic."""
    nfb = 0 if n < 400 else 1 + (int(n) - 400) // 160
    return (nfb + 5) // 6 if nfb > 0 else 0


def make_batches(lens, max_rows, max_segments, capi=None):
    """Length-sorted (ascending, as Audio::CutSplit sorts, audio.cpp:1233-1238) greedy packing into batches of
    at most max_rows packed rows."""
    order = np.argsort(lens, kind="stable")
    batches, cur, rows = [], [], 0
    for i in order:
        T = lfr_frames(int(lens[i]))
        r = T + 1 if T > 0 else 0
        if cur and (rows + r > max_rows or len(cur) >= max_segments):
            batches.append(cur)
            cur, rows = [], 0
        cur.append(int(i))
        rows += r
    if cur:
        batches.append(cur)
    return batches


def workload_config(args, lens, workload):
    """The `config` object: only what defines the workload, so that both arms print the same one."""
    return dict(workload=workload, segments=int(len(lens)), audio_s_per_gpu=float(lens.sum()) / 16000.0,
                max_rows=args.max_rows, batches=len(make_batches(lens, args.max_rows, 4096)),
                l2="inputs+weights per step (>= 790 MB) exceed the 126 MB L2; no explicit flush",
                sharding="each rank runs the full per-GPU workload (weak scaling), no collective")


# ----------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (fp32 PyTorch restatement + C front end), reference harness pattern:
# P worker processes x 1 intra-op thread, each pulling segments from a shared index
# (onnxruntime/bin/funasr-onnx-offline-rtf.cpp:54-102,247-260).
# ----------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_worker(args):
    idx_list, which, keep = args
    import torch
    torch.set_num_threads(1)
    from oracle import frontend as F
    from oracle import paraformer_ref as R
    W = _CPU["Wq"] if which == "int8" else _CPU["W"]
    t0 = time.time()
    outs = []
    for i in idx_list:
        pcm = _CPU["segs"][i]
        feats = F.lfr_cmvn(F.fbank(pcm), _CPU["means"], _CPU["vars"])
        o = R.forward(feats, W, _CPU["pc"], want_taps=False)
        if keep:
            lg = o["logits"].numpy()
            top2 = np.sort(lg, axis=1)[:, -2:] if lg.shape[0] else np.zeros((0, 2), np.float32)
            r = dict(i=i, T=int(feats.shape[0]), alphas=o["alphas"].numpy(), fires=o["fires"].numpy(), ids=np.asarray(o["ids"], np.int32),
                     top_gap=(top2[:, 1] - top2[:, 0]).astype(np.float32), logit_absmax=float(np.abs(lg).max()) if lg.size else 0.0)
            if i < PARITY_TENSORS:
                r["enc"] = o["enc"].numpy()
                r["logits"] = lg
            outs.append(r)
    return time.time() - t0, outs


def sample_indices(lens):
    """The fixed sample: CPU_SAMPLE segments spread evenly over the length-sorted workload (identical on every box)."""
    order = np.argsort(lens, kind="stable")
    return [int(order[k]) for k in np.linspace(0, len(lens) - 1, num=min(len(lens), CPU_SAMPLE)).astype(int)]


def cpu_baseline(cfg, W, means, vars_, seg_pcm16, procs=None, int8=True, keep=True):
    """Times the oracle on the fixed sample (fp32, then the int8 stand-in on the first half of it) and returns its outputs for
    the parity check.  Must run BEFORE CUDA is initialised (fork)."""
    import multiprocessing as mp

    import torch
    from oracle import frontend as F
    from oracle import paraformer_ref as R
    F.lib()
    P = procs or min(os.cpu_count() or 1, 32)
    pc = R.PfConfig.from_dict(cfg)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    # shuffle the (length-sorted) sample with a fixed seed: the interleaved assignment below then gives every worker the same mix
    perm = np.random.default_rng(7).permutation(len(seg_pcm16))
    _CPU.update(W=Wt, pc=pc, means=means, vars=vars_, segs=[seg_pcm16[j].astype(np.float32) / np.float32(32768) for j in perm])
    if int8:
        _CPU["Wq"] = R.quantize_dynamic_int8(Wt)
    ctx = mp.get_context("fork")
    n = len(seg_pcm16)

    def run(which, count, keep_out):
        chunks = [c for c in (list(range(w, count, P)) for w in range(P)) if c]
        t0 = time.time()
        with ctx.Pool(len(chunks)) as pool:
            res = pool.map(_cpu_worker, [(c, which, keep_out) for c in chunks])
        wall = time.time() - t0
        audio = sum(len(_CPU["segs"][i]) for i in range(count)) / 16000.0
        max_thread = max(r[0] for r in res)   # reference: total_time = max over threads of summed inference time
        outs = [None] * len(seg_pcm16)
        for r in res:
            for o in r[1]:
                outs[int(perm[o["i"]])] = o     # back to the caller's order
        return audio / max_thread, len(chunks), audio, wall, outs

    v32, cores, audio32, wall32, outs = run("fp32", n, keep)
    cb = dict(value=v32, unit=UNIT, cores=cores, kind="port", rtfx_per_core=v32 / cores,
              sample="fixed sample: %d segments (%.0f audio-s) spread evenly over the length-sorted workload; fp32 PyTorch restatement + C "
                     "front end, %d procs x 1 thread (decoder-thread-num = %d, intra-op 1), wall %.1fs" % (n, audio32, cores, cores, wall32))
    if int8:
        n8 = max(1, n // 2)
        v8, c8, audio8, wall8, _ = run("int8", n8, False)
        cb["int8"] = dict(value=v8, unit=UNIT, cores=c8, rtfx_per_core=v8 / c8, kind="stand-in",
                          sample="%d segments of the same sample (%.0f audio-s), wall %.1fs; STAND-IN for the reference's deployed "
                                 "model_quant.onnx: PyTorch dynamic int8 Linear on every projection of the port (onnxruntime and the "
                                 "exported graph are absent here), same harness" % (n8, audio8, wall8))
    return cb, outs


# ----------------------------------------------------------------------------------------------------
def sample_clocks(stop, out):
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    _, local, _ = rank_env()
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def clocks_summary(samples):
    if not samples:
        return dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
    sm = sorted(float(s[0]) for s in samples if s[0].replace(".", "").isdigit())
    reasons = set()
    for s in samples:
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
            if v.lower().startswith("active"):
                reasons.add(name)
    return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=float(samples[0][1]) if samples[0][1].replace(".", "").isdigit() else None,
                reasons=sorted(reasons))


def make_stream(seconds=3600.0, seed=4242):
    """configs[3]: speech bursts U[2,20] s separated by silences U[0.3,1.0] s; boundaries = the generator's ground truth."""
    rng = np.random.default_rng(seed)
    t, segs = 0.0, []
    while True:
        t += float(rng.uniform(0.3, 1.0))
        d = round(float(rng.uniform(2.0, 20.0)) * 100.0) / 100.0
        if t + d > seconds:
            break
        segs.append((int(t * 16000), int((t + d) * 16000)))
        t += d
    return segs


def product_multigpu(capi, synth, model_dir, n_gpus, steps):
    """The product's own multi-GPU path: ONE handle (FunOfflineInit with devices = 0..N-1 -> MultiGpuParaformer: per-GPU worker,
    engine and queue, longest-processing-time-first assignment, no collective) driven from this process over all N GPUs with
    host int16 PCM in and text out.  Strong scaling: the job is fixed, N grows."""
    out = {}
    h = capi.OfflineHandle(model_dir, max_rows=65536, max_segments=4096, batch_size=4096, devices=list(range(n_gpus)))
    # configs[3]: 1 h stream, VAD segments in arrival order
    segs = make_stream()
    pcm = np.zeros(3600 * 16000, np.int16)
    blk = None
    for k, (b, e) in enumerate(segs):
        if k % 16 == 0:
            blk = synth.make_audio(21 * 16000 * 16, 77 + k)
        o = (k % 16) * 21 * 16000
        pcm[b:e] = blk[o:o + (e - b)]
    sb, se = [s[0] for s in segs], [s[1] for s in segs]
    h.infer_segments(pcm, sb[:32], se[:32])
    h.infer_segments(pcm, sb, se)
    t0 = time.perf_counter()
    for _ in range(steps):
        text = h.infer_segments(pcm, sb, se)
    dt = (time.perf_counter() - t0) / steps
    out["config4"] = dict(workload="configs[3]: 1 h synthetic stream, %d ground-truth VAD segments (%.0f s of speech) in arrival order through "
                                   "FunOfflineInferSegmentsB200 on one handle" % (len(segs), sum(e - b for b, e in segs) / 16000.0),
                          n_gpus=n_gpus, wall_s=dt, value=3600.0 / dt, unit="stream_s/s", chars=len(text), scaling="strong",
                          segments_per_gpu=h.segments_per_device())
    if n_gpus > 1:
        # self-check of the sharded path on hardware: the same call on a ONE-GPU handle must return the very same text (segments are
        # independent, the engine is batch invariant, and the pool builds the text in the caller's order through one detokeniser, so
        # sharding may not change a character).  The call is made twice: like the reference's Vocab, the detokeniser carries a
        # leading-space decision over from the previous call, so both handles must come from the same previous call.
        h1 = capi.OfflineHandle(model_dir, max_rows=65536, max_segments=4096, batch_size=4096, devices=[0])
        h1.infer_segments(pcm, sb, se)
        text1 = h1.infer_segments(pcm, sb, se)
        h1.close()
        out["config4"]["text_equals_1gpu_handle"] = bool(text1 == text)
    del pcm
    # configs[4]: 256 segments x 60 s (T = 1000 LFR frames)
    one = synth.make_audio(960000, 4321)
    pcm5 = np.tile(one, 256)
    b5 = [i * 960000 for i in range(256)]
    e5 = [(i + 1) * 960000 for i in range(256)]
    h.infer_segments(pcm5, b5, e5)
    t0 = time.perf_counter()
    for _ in range(steps):
        text = h.infer_segments(pcm5, b5, e5)
    dt = (time.perf_counter() - t0) / steps
    out["config5"] = dict(workload="configs[4]: 256 segments x 60 s (T = 1000) through FunOfflineInferSegmentsB200 on one handle",
                          n_gpus=n_gpus, wall_s=dt, value=256 * 60.0 / dt, unit=UNIT, chars=len(text), scaling="strong",
                          segments_per_gpu=h.segments_per_device())
    h.close()
    return out


def config3_leg(capi, synth, steps, n_segments=256, max_rows=65536):
    """BASELINE.json configs[2]: contextual Paraformer + 100 hotwords + timestamp head (random-init full-size weights) on the first
    `n_segments` segments of the configs[1] workload, device-resident PCM, one GPU.  Times what a forward with hotwords and
    timestamps costs (bias decoder, upsampling + BiLSTM head, us_alphas / us_peaks out) next to the hotword compiler."""
    import tempfile
    import torch
    mf = importlib.import_module("asr-2pass_b200.modelfile")
    pcm, offs = synth.make_segments(1024)
    lens = synth.segment_lengths(1024)[:n_segments]
    cfg, W = synth.make_weights(dict(timestamp=1, contextual=1))
    means, vars_ = synth.make_cmvn()
    tmp = tempfile.mkdtemp(prefix="b200pf_c3_")
    mf.write_model_dir(tmp, cfg, W, means, vars_, synth.make_tokens(int(cfg["vocab"])))
    del W
    eng = capi.Engine(tmp, max_rows=max_rows, max_segments=4096)
    rng = np.random.default_rng(0)
    ids = np.zeros((101, 10), np.int32)
    ln = np.zeros(101, np.int32)
    for j in range(100):
        L = int(rng.integers(2, 7))
        ids[j, :L] = rng.integers(3, 8403, L)
        ln[j] = L
    ids[100, 0], ln[100] = 1, 1
    for _ in range(2):
        eng.hotword_embed(ids, ln)
    t0 = time.perf_counter()               # synchronous call (ids in, [n_hw, 512] floats out): wall time of the whole compile
    for _ in range(3):
        hw = eng.hotword_embed(ids, ln)
    hw_ms = (time.perf_counter() - t0) / 3 * 1e3
    groups = make_batches(lens, max_rows, 4096, capi)
    batches = []
    for g in groups:
        buf = np.concatenate([pcm[offs[i]:offs[i + 1]] for i in g])
        ho = np.concatenate([[0], np.cumsum([lens[i] for i in g])]).astype(np.int64)
        b = capi.Batch(eng, len(buf) + 64)
        b.set_hotwords(hw)
        b.stage_s16(buf, ho)
        batches.append(b)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(eng.stream)
    for _ in range(3):
        for b in batches:
            b.run()
    res = [b.collect() for b in batches]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record()
        for _ in range(steps):
            for b in batches:
                b.run()
        ev1.record()
    ev1.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    audio = float(lens.sum()) / 16000.0
    out = dict(workload="configs[2]: contextual Paraformer + 100 hotwords + timestamps, first %d segments of the configs[1] workload "
                        "(%.0f audio-s), random-init full-size weights, device-resident PCM" % (n_segments, audio),
               ms_per_step=ms, value=audio / (ms / 1e3), unit=UNIT, tokens=int(sum(r["n_tokens"] for r in res)),
               timestamp_peaks=int(sum(int((np.asarray(r.get("us_peaks", [])) >= 1.0 - 1e-4).sum()) for r in res)),
               launches=int(sum(b.launches for b in batches)), hotword_compile_ms=hw_ms)
    for b in batches:
        b.close()
    eng.close()
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return out


def main():
    # Exactly ONE line may reach stdout.  Libraries (NCCL's version banner, torchrun notices) print there too,
    # so fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--segments", type=int, default=1024)
    ap.add_argument("--max-rows", type=int, default=int(os.environ.get("B200PF_MAX_ROWS", "196608")))
    ap.add_argument("--prec", default=None, choices=["fp16", "bf16"], help="operand format (default: the library's, fp16)")
    ap.add_argument("--workload", default="config2", choices=["config2", "config5"],
                    help="config2 (default, the bench line): BASELINE.json configs[1]; config5: 32 segments x 60 s per GPU "
                         "(configs[4]: 256 x 60 s over 8 GPUs, T = 1000 LFR frames) - an extra measurement, not the headline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=0.0, help="ignored (kept for old command lines): the CPU sample is fixed")
    ap.add_argument("--no-multigpu-product", action="store_true")
    args = ap.parse_args()
    rank, local, world = rank_env()
    synth = importlib.import_module("asr-2pass_b200.synth")

    # ---- workload (identical on every rank) ----
    lens = synth.segment_lengths(args.segments)
    cfg, W = synth.make_weights()
    means, vars_ = synth.make_cmvn()
    workload = "configs[1]: %d synthetic VAD segments U[2,20] s, length-bucketed, Paraformer-large random-init" % args.segments
    if args.workload == "config5":
        args.segments = 32
        lens = np.full(32, 960000, np.int64)
        workload = "configs[4]: max-length stress, 32 segments x 60 s per GPU (256 over 8 GPUs), T = 1000 LFR frames"
    config = workload_config(args, lens, workload)

    if args.impl == "reference":
        # The reference's own CPU implementation cannot be built or installed here (its neural graph lives in an
        # external model.onnx executed by a stripped libonnxruntime), so this arm times the oracle
        # port of the same path on the host cores, as the tier framing prescribes: every step = the fixed sample.
        if rank != 0:
            return
        pcm, offs = synth.make_segments(args.segments)
        idx = sample_indices(lens)
        segs = [pcm[offs[i]:offs[i + 1]] for i in idx]
        vals = []
        for step in range(args.warmup + args.steps):
            cb, _ = cpu_baseline(cfg, W, means, vars_, segs, int8=(step == args.warmup + args.steps - 1), keep=False)
            if step >= args.warmup:
                vals.append(cb)
        v = float(np.mean([c["value"] for c in vals]))
        cb = vals[-1]
        cb["value"] = v
        cb["rtfx_per_core"] = v / cb["cores"]
        emit(dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                  ms_per_step=None, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                  data="synthetic", impl="reference", config=config, cpu_baseline=cb,
                  e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0)))
        return

    pcm, offs = synth.make_segments(args.segments)
    if args.workload == "config5":
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        one = synth.make_audio(960000, 4321)
        pcm = np.tile(one, 32)
        args.no_cpu_baseline = True
    cb, oracle_outs, sample = None, None, None
    if rank == 0 and world == 1 and args.gpus == 1 and not args.no_cpu_baseline:
        sample = sample_indices(lens)
        cb, oracle_outs = cpu_baseline(cfg, W, means, vars_, [pcm[offs[i]:offs[i + 1]] for i in sample])  # before CUDA init (fork)

    import torch
    import torch.distributed as dist
    capi = importlib.import_module("asr-2pass_b200.capi")
    if capi.device_count() < 1:
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")

    # weak scaling: every rank processes the full per-GPU workload (its own copy of the 1024 segments,
    # differently seeded would change nothing for timing); whole-job audio = world x per-rank audio.
    tmp = tempfile.mkdtemp(prefix="b200pf_bench_%d_" % rank)
    mf = importlib.import_module("asr-2pass_b200.modelfile")
    mf.write_model_dir(tmp, cfg, W, means, vars_, synth.make_tokens(int(cfg["vocab"])))
    del W
    eng = capi.Engine(tmp, device=local, max_rows=args.max_rows, max_segments=4096, prec=args.prec)
    dtype = "fp16" if eng.cfg.precision == capi.PREC_FP16 else "bf16"
    if os.environ.get("B200PF_OVERLAP"):
        eng.set_option("overlap", int(os.environ["B200PF_OVERLAP"]))
    groups = make_batches(lens, args.max_rows, 4096)
    audio_s = float(lens.sum()) / 16000.0

    # per-batch contiguous pinned host PCM (length-sorted order) and device-resident copies
    host_pcm, host_offs, batches = [], [], []
    for g in groups:
        n = int(sum(lens[i] for i in g))
        hp = torch.empty(n, dtype=torch.int16).pin_memory()
        ho = np.zeros(len(g) + 1, np.int64)
        pos = 0
        hv = hp.numpy()
        for k, i in enumerate(g):
            hv[pos:pos + lens[i]] = pcm[offs[i]:offs[i + 1]]
            pos += int(lens[i])
            ho[k + 1] = pos
        host_pcm.append(hp)
        host_offs.append(ho)
        batches.append(capi.Batch(eng, n + 64))
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))
    copy_stream = torch.cuda.Stream(device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: device-resident (`value`) ----
    for b, hp, ho in zip(batches, host_pcm, host_offs):
        b.stage_s16(hp.data_ptr(), ho)
    torch.cuda.synchronize()

    def step_resident():
        for b in batches:
            b.run()

    results = None
    for _ in range(args.warmup):
        step_resident()
    results = [b.collect() for b in batches]
    launches = sum(b.launches for b in batches)
    flops = sum(b.flops for b in batches)
    barrier()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=sample_clocks, args=(stop, samples), daemon=True)
    if rank == 0:
        th.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record()
        for _ in range(args.steps):
            step_resident()
        ev1.record()
    ev1.synchronize()
    barrier()
    ms_resident = ev0.elapsed_time(ev1) / args.steps

    # ---- per-kernel CUDA-event timing (same launches, events on the launching stream) ----
    eng.set_option("profile", 1)
    eng.profile_read(reset=True)
    for _ in range(args.steps):
        step_resident()
    prof = eng.profile_read(reset=True)
    eng.set_option("profile", 0)
    barrier()

    # ---- leg 2: end to end through the C ABI with host buffers (`e2e`) ----
    # second set of batch objects (device PCM + layout + result buffers) so that consecutive steps ping-pong: the next item of
    # the stream is always staged into an object whose previous results were already collected
    batches_b = [capi.Batch(eng, hp.numel() + 64) for hp in host_pcm]
    sets = (batches, batches_b)

    def run_e2e(n_steps):
        # double-buffered across batches AND steps: while one batch computes, the NEXT batch of the stream (the next step's first
        # batch after the last one) is staged on the copy stream; collect = D2H + sync.  Every step's host->device copies and
        # result reads happen inside this call.
        seq = [(s, i) for s in range(n_steps) for i in range(len(batches))]
        sets[0][0].stage_s16(host_pcm[0].data_ptr(), host_offs[0], stream=copy_stream.cuda_stream)
        for k, (s_, i) in enumerate(seq):
            if k + 1 < len(seq):
                s2, j = seq[k + 1]
                sets[s2 & 1][j].stage_s16(host_pcm[j].data_ptr(), host_offs[j], stream=copy_stream.cuda_stream)
            sets[s_ & 1][i].run()
            sets[s_ & 1][i].collect()

    run_e2e(max(1, args.warmup - 1))
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / args.steps
    barrier()
    stop.set()

    h2d = int(sum(hp.numel() * 2 for hp in host_pcm))
    n_tok = int(sum(r["n_tokens"] for r in results))
    d2h = int(sum((2 * len(g) + 1) * 4 for g in groups) + 2 * 4 * sum(sum(lfr_frames(int(lens[i])) + 1 for i in g) for g in groups))

    # ---- leg 3 (rank 0, N = 1): the reference-facing Model::Forward(float**, int*) of the host library: pageable float in, text out ----
    mf_leg = None
    if rank == 0 and world == 1:
        for b in batches_b:
            b.close()
        h = capi.OfflineHandle(tmp, device=local, max_rows=args.max_rows, max_segments=4096, batch_size=4096)
        order = np.argsort(lens, kind="stable")      # the reference sorts the VAD segments by length before Forward (audio.cpp:1233-1238)
        fsegs = [pcm[offs[i]:offs[i + 1]].astype(np.float32) / np.float32(32768) for i in order]
        h.model_forward(fsegs[:64])
        h.model_forward(fsegs)
        ms_mf, n_str = h.model_forward_timed(fsegs, iters=args.steps)    # timed inside the host library, around the virtual call
        dt = ms_mf / 1e3
        mf_leg = dict(value=audio_s / dt, unit=UNIT, ms_per_step=dt * 1e3, h2d_bytes_per_step=int(sum(len(s) * 2 for s in fsegs)),
                      d2h_bytes_per_step=d2h, strings=int(n_str),
                      api="funasr::Model::Forward(float** din, int* len, ...) on one handle: pageable float host buffers in (converted "
                          "exactly to int16 in pinned memory by host threads, then copied), text out")
        h.close()
        del fsegs

    # ---- parity of the timed path on the cpu_baseline sample (rank 0, N = 1) ----
    parity = None
    if oracle_outs is not None:
        P = importlib.import_module("asr-2pass_b200.parity")
        engp = capi.Engine(tmp, device=local, max_rows=32768, max_segments=256, prec=args.prec)
        engp.set_option("taps", 1)
        segs = [pcm[offs[i]:offs[i + 1]] for i in sample]
        so = np.concatenate([[0], np.cumsum([len(s) for s in segs])]).astype(np.int64)
        bp = capi.Batch(engp, int(so[-1]) + 64)
        rp = bp.forward_s16(np.concatenate(segs), so)
        stats = P.new_stats()
        for k, o in enumerate(oracle_outs):
            s, e = rp["token_offsets"][k], rp["token_offsets"][k + 1]
            P.compare_segment(o, int(rp["lfr_frames"][k]), rp["token_ids"][s:e], rp["fire_frames"][s:e], stats,
                              enc=bp.tap("enc", k) if "enc" in o else None, alphas=bp.tap("alphas", k),
                              logits=bp.tap("logits", k) if "logits" in o else None, logit_tol=1e-2, strict=False)
        parity = P.summarize(stats)
        parity["tolerances"] = dict(enc_rel=1e-2, logit_rel=1e-2, note="ids may differ only where the oracle's top-1 margin over the GPU's pick is "
                                    "<= 2e-2 * max|logit|; a fire may move one frame only inside the accumulated alpha deviation")
        parity["reference"] = ("fp32 oracle (oracle/paraformer_ref.py) on the cpu_baseline sample; encoder output and logits compared on "
                               "%d segments of it" % stats["logit_segments"])
        # and the ids of the TIMED batches equal the parity batch's (the engine is batch invariant)
        pos = {}
        for g, r in zip(groups, results):
            for k, i in enumerate(g):
                pos[i] = r["token_ids"][r["token_offsets"][k]:r["token_offsets"][k + 1]]
        parity["timed_batches_equal_parity_batch"] = bool(all(np.array_equal(pos[i], rp["token_ids"][rp["token_offsets"][k]:rp["token_offsets"][k + 1]])
                                                              for k, i in enumerate(sample)))
        bp.close()
        engp.close()

    t = torch.tensor([ms_resident, t_e2e * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_resident, ms_e2e = float(t[0]), float(t[1])

    # ---- the product's multi-GPU path, driven by rank 0 over all N GPUs while the other ranks wait on the CPU ----
    for b in batches:
        b.close()
    eng.close()
    torch.cuda.synchronize()
    multi = None
    if world > 1:
        dist.barrier(group=cpu_group)         # every rank has released its engine; waiting ranks hold no GPU work from here on
    if rank == 0 and not args.no_multigpu_product and args.workload == "config2":
        try:
            multi = product_multigpu(capi, synth, tmp, world, max(2, min(args.steps, 3)))
        except Exception as ex:   # reported, never hidden
            multi = dict(multigpu_product_error=str(ex))
        try:
            multi = dict(multi or {}, config3=config3_leg(capi, synth, max(2, min(args.steps, 3))))
        except Exception as ex:
            multi = dict(multi or {}, config3=dict(error=str(ex)))
    if world > 1:
        dist.barrier(group=cpu_group)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, tflops_peak, which = load_peaks()
    traffic = None   # DRAM bytes per GEMM launch from the committed ncu --set full capture (tools/make_traffic.py)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic = tj["gemm_tcgen05_kernel"]["dram_bytes_per_launch"]
    except Exception:
        pass
    value = world * audio_s / (ms_resident / 1e3)
    e2e_v = world * audio_s / (ms_e2e / 1e3)
    step_tf = flops / (ms_resident / 1e3) / 1e12
    gk = [k for k in prof if k.startswith("gemm_")]
    g = dict(ms=sum(prof[k]["ms"] for k in gk), work=sum(prof[k]["work"] for k in gk), launches=sum(prof[k]["launches"] for k in gk))
    gemm_tf = g["work"] / max(g["ms"], 1e-9) / 1e9 if g["launches"] else 0.0
    prof_total = sum(v["ms"] for v in prof.values()) or 1.0
    kernels = {k: dict(ms_per_step=v["ms"] / args.steps, share=v["ms"] / prof_total, launches_per_step=v["launches"] // args.steps,
                       achieved=(v["work"] / max(v["ms"], 1e-9) / 1e9 if (k.startswith("gemm_") or k == "attention_tcgen05")
                                 else v["work"] / max(v["ms"], 1e-9) / 1e6),
                       unit=("TFLOP/s" if (k.startswith("gemm_") or k == "attention_tcgen05") else "GB/s"))
               for k, v in prof.items() if v["launches"]}
    for k, v in kernels.items():   # fraction of the roofline that bounds the kernel class (measured peaks, MEASURED_PEAKS.json)
        v["frac"] = v["achieved"] / (tflops_peak if v["unit"] == "TFLOP/s" else hbm)
    out = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms_resident, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=dtype,
        dtype_note="16-bit tensor-core operands (tcgen05 kind::f16: IEEE fp16 by default, bf16 with --prec bf16 -- same rate), fp32 "
                   "accumulation, residual streams, LayerNorm statistics, softmax and CIF",
        data="synthetic", config=config, tokens_per_step=n_tok,
        e2e=dict(value=e2e_v, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=ms_e2e),
        gpu_launches=int(launches * args.steps),
        roofline=dict(bound="tensor", kernel="gemm_tcgen05_kernel (all %d GEMM launches of a step)" % (g["launches"] // args.steps), achieved=gemm_tf, peak=tflops_peak, unit="TFLOP/s",
                      frac=gemm_tf / tflops_peak, traffic=traffic,
                      note="sum of 2*M*N*K over the GEMM launches of a step / their CUDA-event time, vs %s sustained 16-bit dense peak; "
                           "whole step: %.1f TFLOP/s (%.3f of peak)" % (which, step_tf, step_tf / tflops_peak)),
        kernels=kernels,
        clocks=clocks_summary(samples),
    )
    if mf_leg is not None:
        out["e2e_model_forward"] = mf_leg
    if parity is not None:
        out["parity"] = parity
    if cb is not None:
        out["cpu_baseline"] = cb
    if multi is not None:
        out.update(multi)
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
