"""CUDA path against what the reference's OWN compiled Paraformer::Forward produced (tests/golden/am_forward_golden.*, generated
by tests/golden/make_golden.py from oracle/am_ref.py: reference host code over a stand-in onnxruntime, network = the fp32 oracle).

  * features: the product's LFR + CMVN tap == the tensor the reference handed to its session, <= 1e-4 (fp32 fbank, double FFT);
  * token ids: equal to the reference run's ids except where the fp32 top-1 margin is below 0.06 (twice the 1e-2 logit tolerance at max|logit| ~ 3; fp16 operands; the
    same rule as tests/test_gpu_parity.py) -- and where ids are equal the result STRING (text, and text | stamps for the
    timestamp model) must equal the reference's byte for byte."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_forward_against_reference_compiled_golden(capi, synth, gpu, tmp_path):
    g = json.load(open(os.path.join(HERE, "golden", "am_forward_golden.json"), encoding="utf-8"))
    arr = np.load(os.path.join(HERE, "golden", "am_forward_golden.npz"))
    models, exact, feats_checked = {}, 0, 0
    per_model_index = {}
    for case in g["forward"]:
        name = case["model"]
        k = per_model_index.get(name, 0)
        per_model_index[name] = k + 1
        if name not in models:
            d = str(tmp_path / name)
            os.makedirs(d)
            synth.write_synthetic_model_dir(d, case["cfg"], seed=case["model_seed"], jitter_ln=True)
            sd = os.path.join(d, "seg_dict")
            with open(sd, "w", encoding="utf-8") as f:
                f.write(g["seg_dict"])
            h = capi.OfflineHandle(d, max_rows=4096, max_segments=64, batch_size=1)
            eng = capi.Engine(d, max_rows=4096, max_segments=64)
            eng.set_option("taps", 1)
            hw = None
            if case["hotwords"]:
                h.init_seg_dict(sd)
                hw = h.compile_hotwords(case["hotwords"])
            models[name] = (h, eng, hw)
        h, eng, hw = models[name]
        pcm = synth.make_audio(case["n_samples"], case["audio_seed"])
        x = pcm.astype(np.float32) / np.float32(32768)
        texts = h.model_forward([x], hw_emb=hw)
        text = texts[0] if isinstance(texts, (list, tuple)) else texts
        b = capi.Batch(eng, len(pcm) + 16)
        if hw is not None:
            b.set_hotwords(hw)
        r = b.forward_s16(pcm, np.array([0, len(pcm)], np.int64))
        key = "feats_%s_%d" % (name, k)
        if key in arr.files:
            f = b.tap("feats", 0)
            assert f.shape == arr[key].shape and np.abs(f - arr[key]).max() <= 1e-4
            feats_checked += 1
        ids = [int(i) for i in r["token_ids"]]
        gold = case.get("ids", [])
        if case["n_samples"] < 400:
            assert text == "" and case["text"] == ""
            b.close()
            continue
        if len(ids) == len(gold):
            for j, (a, c) in enumerate(zip(ids, gold)):
                assert a == c or case["gaps"][j] < 0.06, (name, k, j, case["gaps"][j])
        else:
            assert abs(len(ids) - len(gold)) <= 1      # a fire within the alpha drift of the threshold (test_gpu_parity.py)
        if ids == gold:
            # spaces inside the text depend on Vocab's cross-call English-word state (vocab.cpp:177), i.e. on earlier calls' ids
            t_txt, _, t_st = text.partition(" | ")
            g_txt, _, g_st = case["text"].partition(" | ")
            assert t_txt.replace(" ", "") == g_txt.replace(" ", "") and t_st == g_st, (name, k)
            exact += 1
        b.close()
    assert feats_checked >= 3
    assert exact >= 2, "no case decoded to the reference run's ids: %d" % exact
    for h, eng, _ in models.values():
        h.close()
        eng.close()
