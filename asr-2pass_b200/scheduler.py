"""Host-side segment scheduling for the offline path: length bucketing into packed batches and sharding of
batches over the GPUs of one box (independent per-GPU queues, no collective: SURVEY.md §8(e)).

Mirrors what the reference does on the host before `Model::Forward`:
  * `Audio::CutSplit` sorts VAD segments ascending by length and remembers the permutation
    (onnxruntime/src/audio.cpp:1233-1238); results are un-permuted afterwards
    (onnxruntime/src/funasrruntime.cpp:270-279);
  * `Audio::FetchDynamic` forms greedy batches under a padded-audio cap (audio.cpp:1052-1108).
Here batches are packed (no padding), so the cap is on packed rows instead of padded seconds.
"""
import numpy as np


def num_fbank_frames(n):
    return 0 if n < 400 else 1 + (int(n) - 400) // 160


def num_lfr_frames(n):
    f = num_fbank_frames(n)
    return 0 if f <= 0 else (f + 5) // 6


def segment_cost(T, L=None):
    """Algorithmic FLOPs of one segment; L defaults to T/2 (random-init CIF rate)."""
    L = T / 2.0 if L is None else L
    D, F, V = 512.0, 2048.0, 8404.0
    enc = 50 * (2 * T * D * 3 * D + 2 * T * D * D + 4 * T * D * F + 4 * T * T * D) + 2 * T * (560 - 512) * 3 * D
    dec = 16 * (4 * L * D * F + 4 * L * D * D + 4 * T * D * D + 4 * L * T * D) + 4 * L * D * F + 2 * L * D * V
    return enc + 2 * T * D * D * 3 + dec


def plan_batches(n_samples, max_rows, max_segments=4096):
    """Ascending-length greedy packing.  Returns a list of index lists; every index appears exactly once."""
    n_samples = np.asarray(n_samples, dtype=np.int64)
    order = np.argsort(n_samples, kind="stable")
    batches, cur, rows = [], [], 0
    for i in order:
        T = num_lfr_frames(int(n_samples[i]))
        r = T + 1 if T > 0 else 0
        if r > max_rows:
            raise ValueError("segment %d needs %d rows > max_rows %d" % (i, r, max_rows))
        if cur and (rows + r > max_rows or len(cur) >= max_segments):
            batches.append(cur)
            cur, rows = [], 0
        cur.append(int(i))
        rows += r
    if cur:
        batches.append(cur)
    return batches


def shard_batches(batches, n_samples, world):
    """Longest-processing-time-first assignment of batches to `world` GPU queues.
    Returns a list (per rank) of lists of batch indices."""
    cost = [sum(segment_cost(num_lfr_frames(int(n_samples[i]))) for i in b) for b in batches]
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for bi in sorted(range(len(batches)), key=lambda k: -cost[k]):
        r = int(np.argmin(load))
        out[r].append(bi)
        load[r] += cost[bi]
    for q in out:
        q.sort()
    return out


def shard_segments(n_samples, world, rank, max_rows, max_segments=4096):
    """The batches (lists of segment indices) rank `rank` of `world` processes: segments are dealt round-robin
    in ascending-length order so every rank sees the same length mix, then packed per rank."""
    n_samples = np.asarray(n_samples, dtype=np.int64)
    order = np.argsort(n_samples, kind="stable")
    mine = order[rank::world]
    local = plan_batches(n_samples[mine], max_rows, max_segments)
    return [[int(mine[j]) for j in b] for b in local]


def gather_results(per_rank_results, n_total):
    """per_rank_results: list over ranks of {segment_index: value}.  Returns a list in original order
    (the reference's un-permute through index_vector, funasrruntime.cpp:270-279)."""
    out = [None] * n_total
    for d in per_rank_results:
        for k, v in d.items():
            if out[k] is not None:
                raise ValueError("segment %d produced twice" % k)
            out[k] = v
    if any(v is None for v in out):
        raise ValueError("missing results")
    return out
