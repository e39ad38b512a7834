"""ORACLE (test infrastructure) — ctypes access to the C restatement (frontend_ref.c) and, when built,
to the reference's own compiled kaldi-native-fbank (oracle/_ref/libknf_ref.so, see oracle/Makefile).

Reference call sites: Paraformer::FbankKaldi onnxruntime/src/paraformer.cpp:309-323,
Paraformer::LfrCmvn :421-461, FindMax onnxruntime/src/util.cpp:63-74.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None
_f32p = ctypes.POINTER(ctypes.c_float)


def _ptr(a):
    return a.ctypes.data_as(_f32p)


def build(ref=True):
    """Compile the oracle's C restatement and (if /root/reference exists) the reference knf."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    if ref and os.path.isdir("/root/reference/onnxruntime/third_party/kaldi-native-fbank"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libpf_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = ctypes.CDLL(path)
        L.pf_oracle_num_fbank_frames.argtypes = [ctypes.c_int64]
        L.pf_oracle_num_fbank_frames.restype = ctypes.c_int
        L.pf_oracle_num_lfr_frames.argtypes = [ctypes.c_int]
        L.pf_oracle_fbank.argtypes = [_f32p, ctypes.c_int64, _f32p]
        L.pf_oracle_lfr_cmvn.argtypes = [_f32p, ctypes.c_int, _f32p, _f32p, _f32p]
        L.pf_oracle_find_max.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.POINTER(ctypes.c_int)]
        L.pf_oracle_rfft_packed.argtypes = [_f32p, ctypes.c_int, _f32p]
        L.pf_oracle_mel_matrix.argtypes = [_f32p]
        L.pf_oracle_window.argtypes = [_f32p]
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's own knf, or None when oracle/_ref was not built (it never is on a box without
    /root/reference unless the prebuilt .so travelled with the snapshot)."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libknf_ref.so")
        if not os.path.exists(path):
            return None
        R = ctypes.CDLL(path)
        R.knf_ref_fbank.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int]
        R.knf_ref_rfft.argtypes = [_f32p, ctypes.c_int]
        _REF = R
    return _REF


def num_fbank_frames(n):
    return lib().pf_oracle_num_fbank_frames(int(n))


def num_lfr_frames(n_fb):
    return lib().pf_oracle_num_lfr_frames(int(n_fb)) if n_fb > 0 else 0


def fbank(pcm):
    """pcm: float32 in [-1,1) -> [n_fb, 80] float32 (restatement)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    n_fb = num_fbank_frames(len(pcm))
    out = np.zeros((n_fb, 80), np.float32)
    if n_fb:
        lib().pf_oracle_fbank(_ptr(pcm), len(pcm), _ptr(out))
    return out


def fbank_ref(pcm):
    """Same through the reference's compiled knf."""
    R = ref_lib()
    if R is None:
        raise RuntimeError("oracle/_ref/libknf_ref.so not built")
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    cap = max(1, len(pcm) // 160 + 2)
    out = np.zeros((cap, 80), np.float32)
    n = R.knf_ref_fbank(_ptr(pcm), len(pcm), _ptr(out), cap)
    return out[:n].copy()


def lfr_cmvn(fb, means, vars_):
    fb = np.ascontiguousarray(fb, dtype=np.float32)
    n_fb = fb.shape[0]
    T = num_lfr_frames(n_fb)
    out = np.zeros((T, 560), np.float32)
    if T:
        means = np.ascontiguousarray(means, np.float32)
        vars_ = np.ascontiguousarray(vars_, np.float32)
        lib().pf_oracle_lfr_cmvn(_ptr(fb), n_fb, _ptr(means), _ptr(vars_), _ptr(out))
    return out


def find_max(row):
    row = np.ascontiguousarray(row, np.float32)
    v = ctypes.c_float()
    i = ctypes.c_int()
    lib().pf_oracle_find_max(_ptr(row), len(row), ctypes.byref(v), ctypes.byref(i))
    return v.value, i.value


def greedy_ids(logits, n_len):
    """Paraformer::GreedySearch paraformer.cpp:386-395: FindMax over the first n_len rows."""
    return [find_max(logits[i])[1] for i in range(int(n_len))]


def rfft_packed(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(len(x), np.float32)
    lib().pf_oracle_rfft_packed(_ptr(x), len(x), _ptr(out))
    return out


def mel_matrix():
    out = np.zeros((80, 256), np.float32)
    lib().pf_oracle_mel_matrix(_ptr(out))
    return out


def window():
    out = np.zeros(400, np.float32)
    lib().pf_oracle_window(_ptr(out))
    return out
