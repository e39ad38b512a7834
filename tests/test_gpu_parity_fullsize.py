"""Full-size parity (-m gpu): the NAMED architecture (50 + 16 layers, 215.8 M parameters, random init) on segments drawn
across BASELINE.json configs[1]'s whole length range plus two 60 s segments (configs[4], T = 1000), CUDA path vs the fp32
oracle (oracle/paraformer_ref.py, run in a process pool: tests/oracle_pool.py).

Acceptance, as north_star states it (tolerances written here, used below):
  * encoder output and logits: max-abs error relative to the tensor's max-abs <= 1e-2            (ENC_TOL, LOGIT_TOL)
  * token count: equal, unless the oracle's own sum of alphas is within the accumulated alpha deviation of an integer
  * CIF fire frames: bit-exact wherever the oracle's integrate value is farther from the threshold than the accumulated alpha
    deviation; elsewhere a fire may move by one frame.  At most FIRE_RATE of all fires may move.
  * token ids: bit-exact wherever the oracle's top-1 margin over the GPU's pick exceeds 2 * LOGIT_TOL * max|logit| (the GPU
    may only pick another id where the two logits tie within the stated tolerance); at most ID_RATE of all tokens may differ.
The same statistics are printed by bench.py (`parity` object) on its cpu_baseline sample.
"""
import importlib

import numpy as np
import pytest

from oracle_pool import oracle_forward_many

pytestmark = pytest.mark.gpu

ENC_TOL = 1e-2
LOGIT_TOL = 1e-2
ID_RATE = 0.03     # measured: ~1 % of tokens sit at an fp32 tie narrower than the tolerance (random-init logits, top-2 gap median 0.07)
FIRE_RATE = 0.01


P = importlib.import_module("asr-2pass_b200.parity")
rel = P.rel


def test_full_size_model_on_configs1_sample_and_max_length_segments(capi, synth, gpu, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("fullparity"))
    synth.write_synthetic_model_dir(d, None, seed=0)
    lens_all = synth.segment_lengths(1024)
    pick = np.argsort(lens_all, kind="stable")[np.linspace(0, 1023, 64).astype(int)]     # 64 segments across the 2-20 s range
    lens = [int(lens_all[i]) for i in pick] + [960000, 960000]                             # + 2 x 60 s: T = 1000 (configs[4])
    segs = [synth.make_audio(n, 5000 + k) for k, n in enumerate(lens)]
    oracle = oracle_forward_many(segs, None, seed=0)
    eng = capi.Engine(d, max_rows=16384, max_segments=128)
    assert eng.cfg.precision == capi.PREC_FP16        # the default operand format is the one that meets the stated tolerance
    eng.set_option("taps", 1)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    b = capi.Batch(eng, int(offs[-1]) + 64)
    res = b.forward_s16(np.concatenate(segs), offs)
    stats = P.new_stats()
    for i, o in enumerate(oracle):
        s, e = res["token_offsets"][i], res["token_offsets"][i + 1]
        # the GPU scan applied to the GPU's own alphas equals the reference recurrence bit for bit (index work)
        fires_self = b.tap("fires", i)
        assert np.array_equal(res["fire_frames"][s:e], np.where(fires_self >= 1.0)[0])
        P.compare_segment(o, int(res["lfr_frames"][i]), res["token_ids"][s:e], res["fire_frames"][s:e], stats, enc=b.tap("enc", i),
                          alphas=b.tap("alphas", i), logits=b.tap("logits", i), logit_tol=LOGIT_TOL, strict=True)
    print("full-size parity:", P.summarize(stats))
    assert stats["segments"] == 66 and stats["tokens"] > 3000
    assert stats["enc_rel"] <= ENC_TOL and stats["logit_rel"] <= LOGIT_TOL
    assert stats["ids_differ"] <= ID_RATE * stats["tokens"]
    assert stats["fires_moved"] <= max(2, FIRE_RATE * stats["fires"])
    b.close()
    eng.close()


def test_bf16_operand_mode_is_the_looser_one(capi, synth, gpu, tmp_path_factory):
    """Precision 0 (north_star's literal bf16 operands) stays available and works, but only reaches 2-4 % on the logits --
    the reason it is not the default."""
    d = str(tmp_path_factory.mktemp("bf16mode"))
    synth.write_synthetic_model_dir(d, None, seed=0)
    lens = [160000, 52800]
    segs = [synth.make_audio(n, 1234 + k) for k, n in enumerate(lens)]
    oracle = oracle_forward_many(segs, None, seed=0, procs=2)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    errs = {}
    for prec in ("bf16", "fp16"):
        eng = capi.Engine(d, max_rows=2048, max_segments=16, prec=prec)
        assert eng.cfg.precision == (capi.PREC_BF16 if prec == "bf16" else capi.PREC_FP16)
        eng.set_option("taps", 1)
        b = capi.Batch(eng, int(offs[-1]) + 64)
        res = b.forward_s16(np.concatenate(segs), offs)
        errs[prec] = max(rel(b.tap("enc", i), o["enc"]) for i, o in enumerate(oracle))
        assert abs(int(res["token_counts"][0]) - oracle[0]["token_num"]) <= 1
        b.close()
        eng.close()
    assert errs["fp16"] <= 2.5e-3 and errs["bf16"] <= 1.2e-2 and errs["fp16"] < 0.5 * errs["bf16"]


def test_full_size_results_do_not_depend_on_batching(capi, synth, gpu, tmp_path_factory):
    """Batch invariance at FULL size, bit for bit: the same segments as one batch, as batches of 8, reversed, and one at a time --
    the last takes every small-batch path at once (CUDA graph replay, FSMN and cross-attention k/v projection forked onto the side
    stream, 128-wide GEMM tiles), the first none of them.  Ids, fire frames, encoder output, alphas and logits must be identical:
    it is what lets the host library cut, shard and micro-batch calls freely (MultiGpuParaformer, MicroBatcher, RunAll)."""
    d = str(tmp_path_factory.mktemp("fullinv"))
    synth.write_synthetic_model_dir(d, None, seed=0)
    lens_all = synth.segment_lengths(1024)
    lens = [int(x) for x in lens_all[:: 1024 // 24][:24]]
    segs = [synth.make_audio(n, 700 + k) for k, n in enumerate(lens)]
    eng = capi.Engine(d, max_rows=16384, max_segments=64)
    eng.set_option("taps", 1)
    b = capi.Batch(eng, int(sum(lens)) + 64)

    def run(order, group):
        out = {}
        for g0 in range(0, len(order), group):
            idx = order[g0:g0 + group]
            offs = np.concatenate([[0], np.cumsum([lens[i] for i in idx])]).astype(np.int64)
            r = b.forward_s16(np.concatenate([segs[i] for i in idx]), offs)
            for k, i in enumerate(idx):
                s, e = r["token_offsets"][k], r["token_offsets"][k + 1]
                out[i] = (r["token_ids"][s:e].copy(), r["fire_frames"][s:e].copy(), b.tap("enc", k), b.tap("alphas", k), b.tap("logits", k))
        return out

    order = list(range(24))
    ref = run(order, 24)
    for name, got in (("batches of 8", run(order, 8)), ("reversed", run(order[::-1], 24)), ("one at a time", run(order, 1)),
                      ("one at a time, again (graph replay)", run(order, 1))):
        for i in order:
            for x, y in zip(ref[i], got[i]):
                assert x.shape == y.shape and np.array_equal(x, y), (name, i)
    st = eng.graph_stats()
    assert st["replays"] > 0          # the one-at-a-time runs really went through captured graphs
    b.close()
    eng.close()
