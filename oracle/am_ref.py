"""ORACLE (test infrastructure only): the REFERENCE's own funasr::Paraformer and funasr::CTTransformer classes, compiled in
place from /root/reference by `make -C oracle ref` into oracle/_ref/libfunasr_am_ref.so, running over a stand-in onnxruntime
(oracle/fake_ort.cc).  The network behind a session is a Python callable the caller supplies (normally the oracle's own
restatement of the graph, oracle/paraformer_ref.py); fbank, LFR + CMVN, greedy search, timestamps, detokenisation, hotword id
packing and the punctuation mini-sentence logic are the reference's compiled code.  Only available in the build container; the
golden vectors generated from it (tests/golden/am_forward_golden.json, tests/golden/make_golden.py) travel to the GPU box."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libfunasr_am_ref.so")
_lib = None
_keep = []     # registered callbacks must outlive their sessions

FLOAT, INT32, INT64 = 1, 6, 7     # ONNXTensorElementDataType
_NP = {FLOAT: np.float32, INT32: np.int32, INT64: np.int64}
_ORT = {np.dtype(np.float32): FLOAT, np.dtype(np.int32): INT32, np.dtype(np.int64): INT64}


class _Tensor(C.Structure):
    _fields_ = [("type", C.c_int32), ("ndim", C.c_int32), ("shape", C.c_int64 * 8), ("data", C.c_void_p)]


_RUN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(_Tensor))


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        L.fake_ort_register.argtypes = [C.c_char_p, C.c_int, C.c_int, _RUN, C.c_void_p]
        L.fake_ort_register.restype = None
        L.fake_ort_set_output.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_void_p]
        L.ref_am_create.restype = C.c_void_p
        L.ref_am_create.argtypes = [C.c_char_p] * 4
        L.ref_am_destroy.argtypes = [C.c_void_p]
        L.ref_am_init_hw.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_am_init_seg_dict.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_am_forward.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.ref_am_compile_hotwords.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.c_int, C.c_int]
        L.ref_punc_create.restype = C.c_void_p
        L.ref_punc_create.argtypes = [C.c_char_p] * 3
        L.ref_punc_destroy.argtypes = [C.c_void_p]
        L.ref_punc_add.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        L.ref_punc_online_create.restype = C.c_void_p
        L.ref_punc_online_create.argtypes = [C.c_char_p] * 3
        L.ref_punc_online_destroy.argtypes = [C.c_void_p]
        L.ref_punc_online_add.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.ref_vad_create.restype = C.c_void_p
        L.ref_vad_create.argtypes = [C.c_char_p] * 3
        L.ref_vad_destroy.argtypes = [C.c_void_p]
        L.ref_vad_cutsplit.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
        L.ref_offline_init.restype = C.c_void_p
        L.ref_offline_init.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int]
        L.ref_offline_uninit.argtypes = [C.c_void_p]
        L.ref_offline_infer_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                               C.c_char_p, C.c_int]
        _lib = L
    return _lib


def register_network(path_suffix, n_in, n_out, fn):
    """Sessions the reference creates for a model path ending in `path_suffix` evaluate `fn(list of numpy inputs) -> list of
    n_out numpy outputs` (float32 / int32 / int64)."""
    L = lib()

    def run(_user, ctx, n, tensors):
        try:
            ins = []
            for i in range(n):
                t = tensors[i]
                shape = tuple(int(t.shape[k]) for k in range(t.ndim))
                cnt = int(np.prod(shape)) if shape else 1
                buf = (C.c_char * (cnt * np.dtype(_NP[t.type]).itemsize)).from_address(t.data) if cnt else b""
                ins.append(np.frombuffer(buf, dtype=_NP[t.type]).reshape(shape).copy())
            outs = fn(ins)
            assert len(outs) == n_out, (len(outs), n_out)
            for i, o in enumerate(outs):
                o = np.ascontiguousarray(o)
                shp = (C.c_int64 * max(1, o.ndim))(*o.shape)
                if L.fake_ort_set_output(ctx, i, _ORT[o.dtype], o.ndim, shp, C.c_void_p(o.ctypes.data)) != 0:
                    return 2
            return 0
        except Exception as e:   # surfaces as an Ort::Exception inside the reference, which logs it and returns ""
            import traceback
            traceback.print_exc()
            return 1

    cb = _RUN(run)
    _keep.append(cb)
    L.fake_ort_register(path_suffix.encode(), n_in, n_out, cb, None)


class RefParaformer:
    """funasr::Paraformer: InitAsr(am_model, am.mvn, config.yaml, tokens.json) [+ InitHwCompiler / InitSegDict]; Forward."""

    def __init__(self, model_dir, net, n_out=2, hw_net=None, seg_dict=None, tag="0"):
        d = os.path.abspath(model_dir)
        am = os.path.join(d, "model_%s.onnx" % tag)          # never opened: the stand-in session is keyed by this path
        register_network(am, 3 if hw_net is not None else 2, n_out, net)
        self.h = lib().ref_am_create(am.encode(), os.path.join(d, "am.mvn").encode(), os.path.join(d, "config.yaml").encode(),
                                     os.path.join(d, "tokens.json").encode())
        if hw_net is not None:
            eb = os.path.join(d, "model_eb_%s.onnx" % tag)
            register_network(eb, 1, 1, hw_net)
            lib().ref_am_init_hw(self.h, eb.encode())
        if seg_dict is not None:
            lib().ref_am_init_seg_dict(self.h, seg_dict.encode())

    def forward(self, pcm_f32, hw_emb=None):
        x = np.ascontiguousarray(pcm_f32, dtype=np.float32)
        buf = C.create_string_buffer(1 << 20)
        if hw_emb is None:
            hp, n_hw, dim = None, 0, 0
        else:
            hw = np.ascontiguousarray(hw_emb, dtype=np.float32)
            hp, (n_hw, dim) = hw.ctypes.data_as(C.POINTER(C.c_float)), hw.shape
        n = lib().ref_am_forward(self.h, x.ctypes.data_as(C.POINTER(C.c_float)), len(x), hp, n_hw, dim, buf, len(buf))
        assert n >= 0
        return buf.value.decode("utf-8")

    def compile_hotwords(self, hotwords, dim=512, cap=4200):
        out = np.zeros((cap, dim), np.float32)
        n = lib().ref_am_compile_hotwords(self.h, hotwords.encode("utf-8"), out.ctypes.data_as(C.POINTER(C.c_float)), cap, dim)
        assert n >= 0, n
        return out[:n].copy()

    def close(self):
        if self.h:
            lib().ref_am_destroy(self.h)
            self.h = None


class RefPunc:
    """funasr::CTTransformer: InitPunc(punc_model, config.yaml, tokens.json); AddPunc(text, lang)."""

    def __init__(self, model_dir, net, tag="0"):
        d = os.path.abspath(model_dir)
        pm = os.path.join(d, "punc_%s.onnx" % tag)
        register_network(pm, 2, 1, net)
        self.h = lib().ref_punc_create(pm.encode(), os.path.join(d, "config.yaml").encode(), os.path.join(d, "tokens.json").encode())

    def add_punc(self, text, lang="zh-cn"):
        buf = C.create_string_buffer(max(1 << 16, 16 * len(text.encode("utf-8")) + 64))
        n = lib().ref_punc_add(self.h, text.encode("utf-8"), lang.encode(), buf, len(buf))
        assert n >= 0
        return buf.value.decode("utf-8")

    def close(self):
        if self.h:
            lib().ref_punc_destroy(self.h)
            self.h = None


class RefPuncOnline:
    """funasr::CTTransformerOnline: AddPunc(text, cache, lang); the session gets (ids, lengths, vad_mask, sub_masks)."""

    def __init__(self, model_dir, net, tag="0"):
        d = os.path.abspath(model_dir)
        pm = os.path.join(d, "punc_online_%s.onnx" % tag)
        register_network(pm, 4, 1, net)
        self.h = lib().ref_punc_online_create(pm.encode(), os.path.join(d, "config.yaml").encode(), os.path.join(d, "tokens.json").encode())

    def add_punc(self, text, cache, lang="zh-cn"):
        """cache: list of bytes, updated in place.  Returns the punctuated text of this call."""
        raw = text.encode("utf-8")
        cin = b"".join(w + b"\x01" for w in cache)
        cap = max(1 << 16, 16 * (len(raw) + len(cin)) + 64)
        buf, cbuf = C.create_string_buffer(cap), C.create_string_buffer(cap)
        n = lib().ref_punc_online_add(self.h, raw, cin, lang.encode(), buf, cap, cbuf, cap)
        assert n >= 0
        cache[:] = [w for w in cbuf.value.split(b"\x01")[:-1]]
        return buf.value.decode("utf-8", "replace")

    def close(self):
        if self.h:
            lib().ref_punc_online_destroy(self.h)
            self.h = None


VAD_CONFIG_YAML = """frontend: WavFrontendOnline
frontend_conf:
  fs: 16000
  window: hamming
  n_mels: 80
  frame_length: 25
  frame_shift: 10
  dither: 0.0
  lfr_m: 5
  lfr_n: 1
model_conf:
  max_end_silence_time: 800
  max_single_segment_time: 60000
  speech_noise_thres: %.6f
"""


class RefVad:
    """funasr::FsmnVad + FsmnVadOnline driven like Audio::CutSplit (audio.cpp:1172-1226).  vad_dir holds am.mvn; config.yaml
    (the keys FsmnVad::LoadConfigFromYaml reads, fsmn-vad.cpp:21-52) and an empty model file are written here.  The session
    gets (feats [1,T,400], 4 caches [1,128,19,1]) and must return (scores [1,T,248], 4 caches)."""

    def __init__(self, vad_dir, net, speech_noise_thres=0.6, tag="0"):
        d = os.path.abspath(vad_dir)
        vm = os.path.join(d, "vad_%s.onnx" % tag)
        cfg = os.path.join(d, "vad_config_%s.yaml" % tag)
        with open(cfg, "w") as f:
            f.write(VAD_CONFIG_YAML % speech_noise_thres)
        register_network(vm, 5, 5, net)
        self.h = lib().ref_vad_create(vm.encode(), os.path.join(d, "am.mvn").encode(), cfg.encode())

    def cut_split(self, pcm_f32, vad_tail_sil=800, vad_max_len=15000):
        x = np.ascontiguousarray(pcm_f32, dtype=np.float32)
        out = np.zeros((len(x) // 160 + 8, 2), np.int32)
        n = lib().ref_vad_cutsplit(self.h, x.ctypes.data_as(C.POINTER(C.c_float)), len(x), vad_tail_sil, vad_max_len,
                                   out.ctypes.data_as(C.POINTER(C.c_int)), len(out))
        assert n >= 0
        return out[:n].copy()

    def close(self):
        if self.h:
            lib().ref_vad_destroy(self.h)
            self.h = None


class RefOffline:
    """The reference's exported offline API: FunOfflineInit on real model directories (offline-stream.cpp) and
    FunOfflineInferBuffer (funasrruntime.cpp:208-340).  Each directory gets an empty model.onnx (the reference only checks that
    the file exists; the session behind it is the given callable); the VAD directory also gets the config.yaml FsmnVad reads."""

    def __init__(self, am_dir, am_net, am_outputs=2, vad_dir=None, vad_net=None, vad_thres=0.6, punc_dir=None, punc_net=None):
        kv = {"model-dir": os.path.abspath(am_dir)}
        am = os.path.join(kv["model-dir"], "model.onnx")
        open(am, "ab").close()
        register_network(am, 2, am_outputs, am_net)
        if vad_dir is not None:
            kv["vad-dir"] = os.path.abspath(vad_dir)
            vm = os.path.join(kv["vad-dir"], "model.onnx")
            open(vm, "ab").close()
            with open(os.path.join(kv["vad-dir"], "config.yaml"), "w") as f:
                f.write(VAD_CONFIG_YAML % vad_thres)
            register_network(vm, 5, 5, vad_net)
        if punc_dir is not None:
            kv["punc-dir"] = os.path.abspath(punc_dir)
            pm = os.path.join(kv["punc-dir"], "model.onnx")
            open(pm, "ab").close()
            register_network(pm, 2, 1, punc_net)
        keys = (C.c_char_p * len(kv))(*[k.encode() for k in kv])
        vals = (C.c_char_p * len(kv))(*[v.encode() for v in kv.values()])
        self.h = lib().ref_offline_init(keys, vals, len(kv))
        assert self.h

    def infer_buffer(self, pcm16, vad_tail_sil=800, vad_max_len=60000, cap=1 << 22):
        """-> (text, stamp, stamp_sents) of FunASRGetResult / FunASRGetStamp / FunASRGetStampSents."""
        raw = np.ascontiguousarray(pcm16, dtype="<i2")
        t, st, ss = C.create_string_buffer(cap), C.create_string_buffer(cap), C.create_string_buffer(4 * cap)
        n = lib().ref_offline_infer_buffer(self.h, C.c_void_p(raw.ctypes.data), raw.nbytes, vad_tail_sil, vad_max_len, t, cap, st, cap, ss, 4 * cap)
        assert n >= 0, n
        return t.value.decode("utf-8", "replace"), st.value.decode("utf-8"), ss.value.decode("utf-8", "replace")

    def close(self):
        if self.h:
            lib().ref_offline_uninit(self.h)
            self.h = None
