"""Full-size properties (-m gpu): the configs[1] workload (1024 segments of 2-20 s, 11 054 s of audio, the 215.8 M-parameter
architecture) is far beyond what the CPU oracle can check in a test, so this file checks size-independent properties of
the whole path instead: results do not depend on how the segments are batched (two different packings, and segments run
alone), are deterministic, and satisfy the structural invariants of the path (tokens == CIF fires, fire frames strictly
increasing inside [0, T], ids inside the vocabulary)."""
import hashlib
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(capi, eng, pcm, offs, lens, max_rows, bench):
    groups = bench.make_batches(lens, max_rows, 4096, capi)
    ids, fires, counts = {}, {}, {}
    for g in groups:
        buf = np.concatenate([pcm[offs[i]:offs[i + 1]] for i in g])
        ho = np.concatenate([[0], np.cumsum([lens[i] for i in g])]).astype(np.int64)
        b = capi.Batch(eng, len(buf) + 64)
        r = b.forward_s16(buf, ho)
        for k, i in enumerate(g):
            s, e = r["token_offsets"][k], r["token_offsets"][k + 1]
            ids[i], fires[i], counts[i] = r["token_ids"][s:e].copy(), r["fire_frames"][s:e].copy(), int(r["token_counts"][k])
            assert r["lfr_frames"][k] == capi.lib().b200pf_num_lfr_frames(int(lens[i]))
        b.close()
    return ids, fires, counts, len(groups)


def test_config2_full_workload_is_batching_invariant(capi, synth, gpu, tmp_path_factory):
    bench = importlib.import_module("bench")
    d = str(tmp_path_factory.mktemp("full1024"))
    synth.write_synthetic_model_dir(d, None, seed=0)
    eng = capi.Engine(d, max_rows=98304, max_segments=4096)
    pcm, offs = synth.make_segments(1024)
    lens = synth.segment_lengths(1024)
    ids_a, fires_a, counts_a, nb_a = _run(capi, eng, pcm, offs, lens, 98304, bench)      # 2 big batches
    ids_b, fires_b, counts_b, nb_b = _run(capi, eng, pcm, offs, lens, 12288, bench)      # ~16 small ones
    assert nb_a <= 3 and nb_b >= 12

    def digest(ids, fires):
        h = hashlib.sha256()
        for i in range(1024):
            h.update(np.asarray(ids[i], np.int32).tobytes())
            h.update(np.asarray(fires[i], np.int32).tobytes())
        return h.hexdigest()

    assert digest(ids_a, fires_a) == digest(ids_b, fires_b)            # one checksum over all 103 k tokens, two packings
    total = 0
    for i in range(1024):
        T = capi.lib().b200pf_num_lfr_frames(int(lens[i]))
        n = counts_a[i]
        total += n
        assert len(ids_a[i]) == n == len(fires_a[i])
        assert np.all((ids_a[i] >= 0) & (ids_a[i] < 8404))
        if n:
            assert np.all(np.diff(fires_a[i]) > 0) and fires_a[i][0] >= 0 and fires_a[i][-1] <= T
        assert abs(n - T / 2) <= 0.25 * T + 3                            # random-init alphas ~ 0.5 per frame
    assert 90000 < total < 115000
    # a few segments alone
    for i in (0, 17, 500, 1023):
        b = capi.Batch(eng, int(lens[i]) + 64)
        r = b.forward_s16(pcm[offs[i]:offs[i + 1]], np.array([0, lens[i]], np.int64))
        assert np.array_equal(r["token_ids"], ids_a[i]) and np.array_equal(r["fire_frames"], fires_a[i])
        b.close()
    eng.close()
