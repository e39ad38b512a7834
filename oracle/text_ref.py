"""ORACLE (test infrastructure only): ctypes access to the REFERENCE's own host text code, compiled in place from
/root/reference by `make -C oracle ref` into oracle/_ref/libfunasr_text_ref.so (funasr::Vocab::Vector2StringV2 /
Vector2String, funasr::TimestampOnnx, funasr::PostProcess).  Only available in the build container; the golden vectors
generated from it (tests/golden/text_golden.json, tests/golden/make_golden.py) travel to the GPU box instead."""
import ctypes as C
import json
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libfunasr_text_ref.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        L.ref_vocab_create.restype = C.c_void_p
        L.ref_vocab_create.argtypes = [C.c_char_p]
        L.ref_vocab_destroy.argtypes = [C.c_void_p]
        L.ref_vector2string_v2.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_char_p, C.c_char_p, C.c_int]
        L.ref_greedy_with_stamps.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                             C.c_int, C.c_char_p, C.c_int]
        L.ref_online_create.restype = C.c_void_p
        L.ref_online_create.argtypes = [C.c_float, C.c_float]
        L.ref_online_destroy.argtypes = [C.c_void_p]
        L.ref_cif_search.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int]
        L.ref_pos_emb.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int]
        L.ref_e2e_vad_offline.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_int), C.c_int]
        L.ref_e2e_vad_online.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_int), C.c_int]
        L.ref_timestamp_sentence.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def timestamp_sentence(text, stamp):
    """funasr::TimestampSentence (util.cpp:569-637)."""
    buf = C.create_string_buffer(64 * (len(text.encode("utf-8")) + len(stamp)) + 4096)
    n = lib().ref_timestamp_sentence(text.encode("utf-8"), stamp.encode("utf-8"), buf, len(buf))
    assert n >= 0
    return buf.value.decode("utf-8", "replace")


def e2e_vad(sil_prob, max_end_sil=800, max_seg_ms=15000, thres=0.8, chunk_frames=None):
    """funasr::E2EVadModel (e2e-vad.h) on per-frame silence probabilities -> [[start_ms, end_ms], ...].  chunk_frames=None: one
    offline call; otherwise the chunked online form exactly as Audio::CutSplit drives it (audio.cpp:1172-1226)."""
    p = np.ascontiguousarray(sil_prob, dtype=np.float32)
    out = np.zeros((len(p) + 4, 2), np.int32)
    fp, ip = p.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_int))
    if chunk_frames is None:
        n = lib().ref_e2e_vad_offline(fp, len(p), max_end_sil, max_seg_ms, C.c_float(thres), ip, len(out))
    else:
        n = lib().ref_e2e_vad_online(fp, len(p), int(chunk_frames), max_end_sil, max_seg_ms, C.c_float(thres), ip, len(out))
    return out[:n].copy()


def cif_search(hidden, alphas, threshold=1.0, tail=0.45):
    """ParaformerOnline::CifSearch (paraformer-online.cpp:270-345) with chunk_size {0, T, 0} and is_last_chunk: the offline
    predictor's integrate-and-fire incl. the tail frame.  hidden [T,D], alphas [T] -> token frames [L,D]."""
    h = np.ascontiguousarray(hidden, dtype=np.float32)
    a = np.ascontiguousarray(alphas, dtype=np.float32)
    T_, D = h.shape
    out = np.zeros((T_ + 2, D), np.float32)
    po = lib().ref_online_create(C.c_float(threshold), C.c_float(tail))
    n = lib().ref_cif_search(po, h.ctypes.data_as(C.POINTER(C.c_float)), a.ctypes.data_as(C.POINTER(C.c_float)), T_, D,
                             out.ctypes.data_as(C.POINTER(C.c_float)), T_ + 2)
    lib().ref_online_destroy(po)
    return out[:n].copy()


def pos_emb(T_, depth):
    """ParaformerOnline::GetPosEmb (paraformer-online.cpp:240-268) applied to zeros: the sinusoid table for positions 1..T."""
    f = np.zeros((T_, depth), np.float32)
    po = lib().ref_online_create(C.c_float(1.0), C.c_float(0.45))
    lib().ref_pos_emb(po, f.ctypes.data_as(C.POINTER(C.c_float)), T_, depth)
    lib().ref_online_destroy(po)
    return f


class RefVocab:
    """funasr::Vocab loaded from a tokens.json written to a temp file (Vocab::LoadVocabFromJson, vocab.cpp:46-63)."""

    def __init__(self, tokens):
        self._dir = tempfile.mkdtemp(prefix="ref_vocab_")
        p = os.path.join(self._dir, "tokens.json")
        with open(p, "w", encoding="utf-8") as f:
            json.dump(list(tokens), f, ensure_ascii=False)
        self.h = lib().ref_vocab_create(p.encode())

    def vector2string_v2(self, ids, language=""):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        buf = C.create_string_buffer(1 << 18)
        lib().ref_vector2string_v2(self.h, a.ctypes.data_as(C.POINTER(C.c_int)), len(a), language.encode(), buf, len(buf))
        return buf.value.decode("utf-8")

    def greedy_with_stamps(self, ids, us_alphas, us_peaks):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        al = np.ascontiguousarray(us_alphas, dtype=np.float32)
        pk = np.ascontiguousarray(us_peaks, dtype=np.float32)
        buf = C.create_string_buffer(1 << 18)
        lib().ref_greedy_with_stamps(self.h, a.ctypes.data_as(C.POINTER(C.c_int)), len(a), al.ctypes.data_as(C.POINTER(C.c_float)),
                                     pk.ctypes.data_as(C.POINTER(C.c_float)), len(al), buf, len(buf))
        return buf.value.decode("utf-8")
