"""Parity bookkeeping shared by the GPU parity tests and bench.py's `parity` object: compares what the CUDA path returned for
one segment with what a reference run returned for the same segment, under the acceptance rules north_star states.

The reference outputs arrive as plain numpy arrays (the caller ran the oracle); nothing here imports or runs it.

Rules (tolerances are arguments; the callers state them):
  * encoder output and logits: max-abs error relative to the tensor's max-abs;
  * token count: equal, unless the reference's own sum of alphas is within the accumulated alpha deviation of an integer;
  * CIF fire frames: equal wherever the reference's integrate value is farther from the threshold than the accumulated alpha
    deviation; elsewhere a fire may move by ONE frame;
  * token ids (compared when no fire moved, i.e. the decoder saw the same token embeddings): equal wherever the reference's
    top-1 margin over the GPU's pick exceeds 2 * logit_tol * max|logit| -- the GPU may only pick another id where the two logits
    tie within the stated tolerance.
A violated rule is recorded in stats["violations"] (and raised when strict=True); rates are for the caller to bound.
"""
import numpy as np


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float32) - np.asarray(b, np.float32)).max() / (np.abs(b).max() + 1e-30))


def new_stats():
    return dict(segments=0, tokens=0, fires=0, ids_differ=0, fires_moved=0, count_moved=0, enc_rel=0.0, logit_rel=0.0, alpha_abs=0.0,
                logit_segments=0, violations=0)


def compare_segment(ref, T, ids_gpu, fire_frames_gpu, stats, enc=None, alphas=None, logits=None, logit_tol=1e-2, strict=True):
    """ref: dict with T, alphas [T+1], fires [T+1], ids, and optionally enc [T,512], logits [L,V] (fp32 reference run).
    enc / alphas / logits: the CUDA path's taps for the segment (optional: without alphas the fire rule uses a 2e-3 drift floor)."""
    def fail(msg):
        stats["violations"] += 1
        if strict:
            raise AssertionError(msg)

    if T != ref["T"]:
        fail("frame count %d != %d" % (T, ref["T"]))
        return
    stats["segments"] += 1
    al_o = np.asarray(ref["alphas"], np.float32)
    if enc is not None and "enc" in ref:
        stats["enc_rel"] = max(stats["enc_rel"], rel(enc, ref["enc"]))
    if alphas is not None:
        stats["alpha_abs"] = max(stats["alpha_abs"], float(np.abs(alphas - al_o).max()))
        drift = np.abs(np.cumsum(alphas.astype(np.float64)) - np.cumsum(al_o.astype(np.float64)))
    else:
        drift = np.full(len(al_o), 2e-3)
    fires_o = np.asarray(ref["fires"], np.float32)
    fr_o = np.where(fires_o >= 1.0)[0]
    fr = np.asarray(fire_frames_gpu)
    stats["fires"] += len(fr_o)
    s_o = float(al_o.astype(np.float64).sum())
    if len(fr) != len(fr_o):
        if not (abs(len(fr) - len(fr_o)) == 1 and min(s_o - np.floor(s_o), np.ceil(s_o) - s_o) <= drift.max() + 1e-3):
            fail("token count %d vs %d although the reference's alpha sum %.4f is not at an integer" % (len(fr), len(fr_o), s_o))
        stats["count_moved"] += 1
        stats["fires_moved"] += 1
        return
    moved = 0
    for a, c in zip(fr, fr_o):
        if a != c:
            lo, hi = min(a, c), max(a, c)
            margin = min(abs(fires_o[a] - 1.0), abs(fires_o[c] - 1.0))
            if not (hi - lo == 1 and margin <= drift[:hi + 1].max() + 1e-3):
                fail("fire moved %d -> %d with margin %.4f > drift %.4f" % (c, a, margin, drift[:hi + 1].max()))
            moved += 1
    stats["fires_moved"] += moved
    if moved or len(fr) == 0:
        return          # the decoder saw different token embeddings: ids / logits are not comparable row by row
    ids_o = np.asarray(ref["ids"])
    ids_gpu = np.asarray(ids_gpu)
    stats["tokens"] += len(ids_gpu)
    differ = np.where(ids_gpu != ids_o)[0]
    stats["ids_differ"] += len(differ)
    if "logits" in ref:
        lg_o = ref["logits"]
        tol_abs = 2.0 * logit_tol * float(np.abs(lg_o).max())
        for j in differ:
            if lg_o[j, ids_o[j]] - lg_o[j, ids_gpu[j]] > tol_abs:
                fail("row %d: id %d instead of %d although the reference margin is %.4f > %.4f" % (j, ids_gpu[j], ids_o[j], float(lg_o[j, ids_o[j]] - lg_o[j, ids_gpu[j]]), tol_abs))
        if logits is not None:
            stats["logit_segments"] += 1
            stats["logit_rel"] = max(stats["logit_rel"], rel(logits, lg_o))
            if not np.array_equal(ids_gpu, np.asarray(logits).argmax(1)):
                fail("fused argmax differs from the argmax of the GPU's own logits")
    elif "top_gap" in ref:
        tol_abs = 2.0 * logit_tol * float(ref.get("logit_absmax", 3.0))
        for j in differ:
            if ref["top_gap"][j] > tol_abs:
                fail("row %d: id differs although the reference top-2 gap is %.4f > %.4f" % (j, float(ref["top_gap"][j]), tol_abs))


def summarize(stats):
    """The `parity` object bench.py prints."""
    return dict(segments=stats["segments"], tokens=stats["tokens"], fires=stats["fires"],
                token_mismatch_rate=(stats["ids_differ"] / stats["tokens"]) if stats["tokens"] else None,
                fire_moved_rate=(stats["fires_moved"] / stats["fires"]) if stats["fires"] else None,
                token_count_differs=stats["count_moved"], max_enc_rel_err=stats["enc_rel"], max_logit_rel_err=stats["logit_rel"],
                max_alpha_abs_err=stats["alpha_abs"], logit_segments=stats["logit_segments"], rule_violations=stats["violations"])
