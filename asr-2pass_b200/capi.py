"""ctypes binding of the C ABI in include/b200pf.h (libb200pf.so, built in-tree by csrc/Makefile).

This is plumbing for the Python tests and bench.py; it adds no computation.  There is no fallback:
if the shared library is missing, or there is no sm_100 device, calls raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200pf.so")
_lib = None

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_i16p = C.POINTER(C.c_int16)


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("feat_dim", "d_model", "n_heads", "d_ff", "n_enc", "n_dec", "kernel", "vocab", "pred_residual")] + \
               [(n, C.c_float) for n in ("cif_threshold", "tail_threshold", "ln_eps")] + \
               [(n, C.c_int32) for n in ("sample_rate", "max_rows", "max_segments", "timestamp", "contextual", "precision")]


PREC_BF16, PREC_FP16 = 0, 1


class Result(C.Structure):
    _fields_ = [("token_counts", c_i32p), ("token_offsets", c_i32p), ("lfr_frames", c_i32p),
                ("token_ids", c_i32p), ("fire_frames", c_i32p), ("cap_tokens", C.c_int64), ("n_tokens", C.c_int64),
                ("us_alphas", c_f32p), ("us_peaks", c_f32p), ("us_offsets", c_i32p), ("cap_us", C.c_int64),
                ("token_lse", c_f32p), ("topk_logprob", c_f32p), ("topk_ids", c_i32p), ("topk_k", C.c_int32)]


class B200PFError(RuntimeError):
    pass


def build_library(verbose=False):
    """nvcc-compile csrc/ for sm_100a into lib/libb200pf.so (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)


EXPORTS = [
    "b200pf_last_error", "b200pf_version", "b200pf_device_count", "b200pf_model_dir_probe", "b200pf_engine_create", "b200pf_engine_create_prec", "b200pf_op_set_precision", "b200pf_engine_destroy",
    "b200pf_engine_config", "b200pf_engine_vocab_size", "b200pf_engine_token", "b200pf_engine_lang",
    "b200pf_engine_set_option", "b200pf_engine_graph_stats", "b200pf_engine_profile_read", "b200pf_engine_stream", "b200pf_engine_copy_stream", "b200pf_num_fbank_frames", "b200pf_num_lfr_frames",
    "b200pf_rows_for", "b200pf_batch_create", "b200pf_batch_destroy", "b200pf_batch_stage_s16",
    "b200pf_batch_stage_f32", "b200pf_batch_stage_s16_ptrs", "b200pf_batch_run", "b200pf_batch_collect", "b200pf_forward_s16", "b200pf_forward_f32",
    "b200pf_batch_launches", "b200pf_batch_flops", "b200pf_batch_tap", "b200pf_op_gemm", "b200pf_op_gemm_bench", "b200pf_op_conv3",
    "b200pf_op_layernorm", "b200pf_op_attention", "b200pf_op_attention_bench", "b200pf_op_fsmn", "b200pf_op_cif", "b200pf_op_frontend",
    "b200pf_batch_set_hotwords", "b200pf_engine_hotword_embed", "b200pf_op_lstm", "b200pf_op_us_peaks", "b200pf_op_lstm_bench", "b200pf_op_logprob_topk", "b200pf_vad_create", "b200pf_vad_destroy", "b200pf_vad_scores_s16",
    "b200pf_punc_create", "b200pf_punc_destroy", "b200pf_punc_info", "b200pf_punc_infer", "b200pf_punc_infer_vad", "b200pf_punc_launches",
]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200PFError("libb200pf.so is not built (run __graft_entry__.build() or make -C asr-2pass_b200/csrc); "
                          "there is no fallback path")
    L = C.CDLL(LIB_PATH)
    L.b200pf_last_error.restype = C.c_char_p
    L.b200pf_engine_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.b200pf_engine_create_prec.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.b200pf_model_dir_probe.argtypes = [C.c_char_p, C.POINTER(Config), c_i32p, c_i32p]
    L.b200pf_engine_destroy.argtypes = [C.c_void_p]
    L.b200pf_engine_destroy.restype = None
    L.b200pf_engine_config.argtypes = [C.c_void_p, C.POINTER(Config)]
    L.b200pf_engine_vocab_size.argtypes = [C.c_void_p]
    L.b200pf_engine_token.argtypes = [C.c_void_p, C.c_int]
    L.b200pf_engine_token.restype = C.c_char_p
    L.b200pf_engine_lang.argtypes = [C.c_void_p]
    L.b200pf_engine_lang.restype = C.c_char_p
    L.b200pf_engine_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.b200pf_engine_profile_read.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    L.b200pf_engine_stream.argtypes = [C.c_void_p]
    L.b200pf_engine_stream.restype = C.c_void_p
    L.b200pf_num_fbank_frames.argtypes = [C.c_int64]
    L.b200pf_num_lfr_frames.argtypes = [C.c_int64]
    L.b200pf_rows_for.argtypes = [c_i64p, C.c_int]
    L.b200pf_rows_for.restype = C.c_int64
    L.b200pf_batch_create.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
    L.b200pf_batch_destroy.argtypes = [C.c_void_p]
    L.b200pf_batch_destroy.restype = None
    L.b200pf_batch_stage_s16.argtypes = [C.c_void_p, C.c_void_p, c_i64p, C.c_int, C.c_void_p]
    L.b200pf_batch_stage_f32.argtypes = [C.c_void_p, C.POINTER(c_f32p), c_i32p, C.c_int, C.c_void_p]
    L.b200pf_batch_run.argtypes = [C.c_void_p, C.c_void_p]
    L.b200pf_batch_collect.argtypes = [C.c_void_p, C.POINTER(Result), C.c_void_p]
    L.b200pf_forward_s16.argtypes = [C.c_void_p, C.c_void_p, c_i64p, C.c_int, C.POINTER(Result)]
    L.b200pf_forward_f32.argtypes = [C.c_void_p, C.POINTER(c_f32p), c_i32p, C.c_int, C.POINTER(Result)]
    L.b200pf_batch_launches.argtypes = [C.c_void_p]
    L.b200pf_batch_launches.restype = C.c_int64
    L.b200pf_batch_flops.argtypes = [C.c_void_p]
    L.b200pf_batch_flops.restype = C.c_double
    L.b200pf_batch_tap.argtypes = [C.c_void_p, C.c_char_p, C.c_int, c_f32p, C.c_int64, c_i64p]
    L.b200pf_op_gemm.argtypes = [C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, c_f32p, c_i32p]
    L.b200pf_op_gemm_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_f32p]
    L.b200pf_op_conv3.argtypes = [C.c_int, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, c_f32p]
    L.b200pf_op_layernorm.argtypes = [C.c_int, c_f32p, C.c_int, C.c_int, c_f32p, c_f32p, C.c_float, C.c_int, c_f32p, c_f32p]
    L.b200pf_op_attention.argtypes = [C.c_int, c_f32p, c_f32p, c_f32p, c_i32p, c_i32p, c_i32p, c_i32p, C.c_int, C.c_int,
                                      C.c_int64, C.c_int64, C.c_int, c_f32p]
    L.b200pf_op_attention_bench.argtypes = [C.c_int, c_i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_f32p, C.POINTER(C.c_double)]
    L.b200pf_op_fsmn.argtypes = [C.c_int, c_f32p, c_f32p, c_i32p, C.c_int, c_f32p]
    L.b200pf_op_cif.argtypes = [C.c_int, c_f32p, c_f32p, c_i32p, C.c_int, C.c_float, c_i32p, c_f32p, c_f32p, c_i32p, C.c_int64]
    L.b200pf_op_frontend.argtypes = [C.c_void_p, c_i16p, C.c_int64, c_f32p, c_f32p]
    _lib = L
    return L


HOST_LIB_PATH = os.path.join(_HERE, "lib", "libfunasr_b200.so")
_host = None
HOST_EXPORTS = ["b200pf_host_detok_create", "b200pf_host_detok_destroy", "b200pf_host_detok_text", "b200pf_host_detok_text_state", "b200pf_host_timestamp_text",
                "b200pf_host_stitch", "b200pf_host_offline_init", "b200pf_host_offline_uninit", "b200pf_host_offline_infer_buffer",
                "b200pf_host_offline_infer_segments", "b200pf_host_model_forward", "b200pf_host_compile_hotwords",
                "b200pf_host_init_seg_dict", "b200pf_host_model_forward_hw", "b200pf_host_offline_infer_buffer_hw",
                "b200pf_host_mb_create", "b200pf_host_mb_create_mock", "b200pf_host_mb_destroy", "b200pf_host_mb_forward",
                "b200pf_host_mb_stats", "b200pf_host_offline_init_devices", "b200pf_host_partition", "b200pf_host_segments_per_device", "b200pf_host_funasr_infer", "b200pf_host_vad_segments",
                "b200pf_host_offline_init_vad", "b200pf_host_offline_vad_cut", "b200pf_host_offline_infer_buffer_vad", "b200pf_host_pack_hotwords", "b200pf_host_punc_tokenize",
                "b200pf_host_punc_add_scripted", "b200pf_host_punc_create", "b200pf_host_punc_destroy", "b200pf_host_punc_rounds", "b200pf_host_punc_add",
                "b200pf_host_punc_add_batch", "b200pf_host_punc_online_add_scripted", "b200pf_host_punc_online_create",
                "b200pf_host_punc_online_destroy", "b200pf_host_punc_online_add", "b200pf_host_sentence_stamps", "b200pf_host_offline_init_kv",
                "b200pf_host_offline_infer_full", "b200pf_host_tpass_init_kv", "b200pf_host_tpass_online_init", "b200pf_host_tpass_uninit",
                "b200pf_host_tpass_online_uninit", "b200pf_host_tpass_infer", "b200pf_host_vad_segments_streaming", "b200pf_host_expand_posteriors", "b200pf_host_model_forward_timed"]


def host_lib():
    """libfunasr_b200.so: the C++ mirror of funasr::Model / funasrruntime.h (include/b200pf_host.h hooks)."""
    global _host
    if _host is not None:
        return _host
    if not os.path.exists(HOST_LIB_PATH):
        raise B200PFError("libfunasr_b200.so is not built")
    lib()
    H = C.CDLL(HOST_LIB_PATH)
    H.b200pf_host_detok_create.argtypes = [C.POINTER(C.c_char_p), C.c_int]
    H.b200pf_host_detok_create.restype = C.c_void_p
    H.b200pf_host_detok_destroy.argtypes = [C.c_void_p]
    H.b200pf_host_detok_destroy.restype = None
    H.b200pf_host_detok_text.argtypes = [C.c_void_p, c_i32p, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
    H.b200pf_host_timestamp_text.argtypes = [C.c_void_p, c_i32p, C.c_int, c_f32p, c_f32p, C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_stitch.argtypes = [C.POINTER(C.c_char_p), c_f32p, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_offline_init.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    H.b200pf_host_offline_init.restype = C.c_void_p
    H.b200pf_host_offline_uninit.argtypes = [C.c_void_p]
    H.b200pf_host_offline_uninit.restype = None
    H.b200pf_host_offline_infer_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int, c_f32p]
    H.b200pf_host_offline_infer_segments.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, c_i64p, c_i64p, C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_model_forward.argtypes = [C.c_void_p, C.POINTER(c_f32p), c_i32p, C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_compile_hotwords.argtypes = [C.c_void_p, C.c_char_p, c_f32p, C.c_int, C.c_int]
    H.b200pf_host_init_seg_dict.argtypes = [C.c_void_p, C.c_char_p]
    H.b200pf_host_model_forward_hw.argtypes = [C.c_void_p, C.POINTER(c_f32p), c_i32p, C.c_int, c_f32p, C.c_int, C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_offline_infer_buffer_hw.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_f32p, C.c_int, C.c_int, C.c_char_p,
                                                      C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_offline_init_devices.argtypes = [C.c_char_p, c_i32p, C.c_int, C.c_int, C.c_int, C.c_int]
    H.b200pf_host_offline_init_devices.restype = C.c_void_p
    H.b200pf_host_partition.argtypes = [c_i32p, C.c_int, C.c_int, c_i32p]
    H.b200pf_host_segments_per_device.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int]
    H.b200pf_host_funasr_infer.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_vad_segments.argtypes = [c_f32p, C.c_int, C.c_int, C.c_int, C.c_float, c_i32p, C.c_int]
    H.b200pf_host_offline_init_vad.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
    H.b200pf_host_offline_init_vad.restype = C.c_void_p
    H.b200pf_host_offline_vad_cut.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, c_i32p, C.c_int]
    H.b200pf_host_offline_infer_buffer_vad.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_pack_hotwords.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_char_p, c_i32p, c_i32p, C.c_int]
    H.b200pf_host_offline_init_kv.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int, C.c_int]
    H.b200pf_host_offline_init_kv.restype = C.c_void_p
    H.b200pf_host_tpass_init_kv.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int]
    H.b200pf_host_tpass_init_kv.restype = C.c_void_p
    H.b200pf_host_tpass_online_init.argtypes = [C.c_void_p]
    H.b200pf_host_tpass_online_init.restype = C.c_void_p
    H.b200pf_host_tpass_uninit.argtypes = [C.c_void_p]
    H.b200pf_host_tpass_uninit.restype = None
    H.b200pf_host_tpass_online_uninit.argtypes = [C.c_void_p]
    H.b200pf_host_tpass_online_uninit.restype = None
    H.b200pf_host_tpass_infer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                          C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_expand_posteriors.argtypes = [c_f32p, c_i32p, C.c_int, C.c_int, C.c_int, c_f32p]
    H.b200pf_host_vad_segments_streaming.argtypes = [c_f32p, C.c_int, c_i32p, C.c_int, C.c_int, C.c_int, C.c_float, c_i32p, C.c_int]
    H.b200pf_host_offline_infer_full.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                                 C.c_char_p, C.c_int]
    H.b200pf_host_sentence_stamps.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    H.b200pf_host_punc_tokenize.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, c_i32p, C.c_int]
    H.b200pf_host_punc_add_scripted.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_char_p,
                                                C.c_int, C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_punc_create.argtypes = [C.c_char_p, C.c_int, C.c_int]
    H.b200pf_host_punc_create.restype = C.c_void_p
    H.b200pf_host_punc_destroy.argtypes = [C.c_void_p]
    H.b200pf_host_punc_destroy.restype = None
    H.b200pf_host_punc_online_add_scripted.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_char_p,
                                                       C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_punc_online_create.argtypes = [C.c_char_p, C.c_int, C.c_int]
    H.b200pf_host_punc_online_create.restype = C.c_void_p
    H.b200pf_host_punc_online_destroy.argtypes = [C.c_void_p]
    H.b200pf_host_punc_online_destroy.restype = None
    H.b200pf_host_punc_online_add.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_punc_rounds.argtypes = [C.c_void_p]
    H.b200pf_host_punc_rounds.restype = C.c_longlong
    H.b200pf_host_punc_add.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    H.b200pf_host_punc_add_batch.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_char_p, C.c_int, c_i32p]
    H.b200pf_host_mb_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    H.b200pf_host_mb_create.restype = C.c_void_p
    H.b200pf_host_mb_create_mock.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    H.b200pf_host_mb_create_mock.restype = C.c_void_p
    H.b200pf_host_mb_destroy.argtypes = [C.c_void_p]
    H.b200pf_host_mb_destroy.restype = None
    H.b200pf_host_mb_forward.argtypes = [C.c_void_p, c_f32p, C.c_int, c_f32p, C.c_int, C.c_int, C.c_char_p, C.c_int]
    H.b200pf_host_mb_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    _host = H
    return H


class HostDetok:
    """pf::host::Detokenizer (Vocab::Vector2StringV2 / TimestampOnnx / PostProcess mirror)."""

    def __init__(self, tokens):
        arr = (C.c_char_p * len(tokens))(*[t.encode("utf-8") for t in tokens])
        self.h = host_lib().b200pf_host_detok_create(arr, len(tokens))

    def __del__(self):
        try:
            host_lib().b200pf_host_detok_destroy(self.h)
        except Exception:
            pass

    def text(self, ids, lang=""):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        buf = C.create_string_buffer(1 << 16)
        host_lib().b200pf_host_detok_text(self.h, _p(ids, c_i32p), len(ids), lang.encode(), buf, len(buf))
        return buf.value.decode("utf-8")

    def text_state(self, ids, state_in, lang=""):
        """(text, state_out) for an explicit incoming state; the object's own state is not touched."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        buf = C.create_string_buffer(1 << 16)
        so = C.c_int(0)
        H = host_lib()
        H.b200pf_host_detok_text_state.argtypes = [C.c_void_p, c_i32p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.c_char_p, C.c_int]
        H.b200pf_host_detok_text_state(self.h, _p(ids, c_i32p), len(ids), lang.encode(), int(bool(state_in)), C.byref(so), buf, len(buf))
        return buf.value.decode("utf-8"), bool(so.value)

    def timestamp_text(self, ids, us_alphas, us_peaks):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        a, p = _f32(us_alphas), _f32(us_peaks)
        buf = C.create_string_buffer(1 << 18)
        host_lib().b200pf_host_timestamp_text(self.h, _p(ids, c_i32p), len(ids), _p(a), _p(p), len(a), buf, len(buf))
        return buf.value.decode("utf-8")


def host_stitch(msgs, starts, lang):
    arr = (C.c_char_p * len(msgs))(*[m.encode("utf-8") for m in msgs])
    st = _f32(starts)
    t, s = C.create_string_buffer(1 << 18), C.create_string_buffer(1 << 18)
    host_lib().b200pf_host_stitch(arr, _p(st), len(msgs), lang.encode(), t, len(t), s, len(s))
    return t.value.decode("utf-8"), s.value.decode("utf-8")


class OfflineHandle:
    """FunOfflineInit / FunOfflineInferBuffer / FunOfflineUninit through the host shim."""

    def __init__(self, model_dir, device=0, max_rows=0, max_segments=0, batch_size=64, devices=None, vad_dir=None, vad_thres=0.0, options=None):
        """devices=[0, 1, ...]: one engine per listed GPU behind this handle (funasr_b200::MultiGpuParaformer).
        vad_dir: FSMN-VAD model directory; infer_buffer then cuts recordings the way the reference's UseVad() branch does."""
        if options is not None:   # the reference's own key/value map, e.g. {"vad-dir": ..., "punc-dir": ...}
            kv = dict(options)
            kv.setdefault("model-dir", model_dir)
            kv.setdefault("device", str(device))
            kv.setdefault("max-rows", str(max_rows))
            kv.setdefault("max-segments", str(max_segments))
            keys = (C.c_char_p * len(kv))(*[k.encode() for k in kv])
            vals = (C.c_char_p * len(kv))(*[str(v).encode() for v in kv.values()])
            self.h = host_lib().b200pf_host_offline_init_kv(keys, vals, len(kv), batch_size)
        elif vad_dir is not None:
            self.h = host_lib().b200pf_host_offline_init_vad(model_dir.encode(), vad_dir.encode(), device, max_rows, max_segments, batch_size,
                                                             C.c_float(vad_thres))
        elif devices is not None and len(devices) > 1:
            dv = np.ascontiguousarray(devices, dtype=np.int32)
            self.h = host_lib().b200pf_host_offline_init_devices(model_dir.encode(), _p(dv, c_i32p), len(dv), max_rows, max_segments, batch_size)
        else:
            self.h = host_lib().b200pf_host_offline_init(model_dir.encode(), device if not devices else devices[0], max_rows, max_segments, batch_size)
        if not self.h:
            raise B200PFError("FunOfflineInit failed: " + lib().b200pf_last_error().decode("utf-8", "replace"))

    def close(self):
        if self.h:
            host_lib().b200pf_host_offline_uninit(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def infer_buffer(self, pcm16, vad_max_len=60000):
        pcm16 = np.ascontiguousarray(pcm16, dtype="<i2")
        buf = C.create_string_buffer(1 << 20)
        sn = C.c_float()
        n = host_lib().b200pf_host_offline_infer_buffer(self.h, C.c_void_p(pcm16.ctypes.data), pcm16.nbytes, vad_max_len, buf, len(buf),
                                                        C.byref(sn))
        if n < 0:
            raise B200PFError("FunOfflineInferBuffer returned nullptr")
        return buf.value.decode("utf-8"), sn.value

    def vad_cut(self, pcm16, vad_tail_sil=800, vad_max_len=60000):
        """[n, 2] int32 of [start_ms, end_ms): the cut FunOfflineInferBuffer applies when the handle has a VAD model."""
        pcm16 = np.ascontiguousarray(pcm16, dtype=np.int16)
        out = np.zeros((len(pcm16) // 160 + 8, 2), np.int32)
        n = host_lib().b200pf_host_offline_vad_cut(self.h, C.c_void_p(pcm16.ctypes.data), len(pcm16), vad_tail_sil, vad_max_len,
                                                   _p(out, c_i32p), len(out))
        if n < 0:
            raise B200PFError("VAD cut failed: " + lib().b200pf_last_error().decode("utf-8", "replace"))
        return out[:n].copy()

    def infer_buffer_vad(self, pcm16, vad_tail_sil=800, vad_max_len=60000, cap=1 << 22):
        pcm16 = np.ascontiguousarray(pcm16, dtype="<i2")
        buf, st = C.create_string_buffer(cap), C.create_string_buffer(cap)
        n = host_lib().b200pf_host_offline_infer_buffer_vad(self.h, C.c_void_p(pcm16.ctypes.data), pcm16.nbytes, vad_tail_sil, vad_max_len,
                                                            buf, len(buf), st, len(st))
        if n < 0:
            raise B200PFError("FunOfflineInferBuffer returned nullptr")
        return buf.value.decode("utf-8"), st.value.decode("utf-8")

    def infer_full(self, pcm16, vad_tail_sil=800, vad_max_len=60000, cap=1 << 22):
        """FunOfflineInferBuffer -> (text, stamp, stamp_sents)."""
        pcm16 = np.ascontiguousarray(pcm16, dtype="<i2")
        t, st, ss = C.create_string_buffer(cap), C.create_string_buffer(cap), C.create_string_buffer(4 * cap)
        n = host_lib().b200pf_host_offline_infer_full(self.h, C.c_void_p(pcm16.ctypes.data), pcm16.nbytes, vad_tail_sil, vad_max_len,
                                                      t, len(t), st, len(st), ss, len(ss))
        if n < 0:
            raise B200PFError("FunOfflineInferBuffer returned nullptr")
        return t.value.decode("utf-8", "replace"), st.value.decode("utf-8"), ss.value.decode("utf-8", "replace")

    def infer_segments(self, pcm16, seg_begin, seg_end, cap=1 << 22):
        pcm16 = np.ascontiguousarray(pcm16, dtype=np.int16)
        b = np.ascontiguousarray(seg_begin, dtype=np.int64)
        e = np.ascontiguousarray(seg_end, dtype=np.int64)
        buf = C.create_string_buffer(cap)
        n = host_lib().b200pf_host_offline_infer_segments(self.h, C.c_void_p(pcm16.ctypes.data), len(pcm16), _p(b, c_i64p), _p(e, c_i64p),
                                                          len(b), buf, len(buf))
        if n < 0:
            raise B200PFError("FunOfflineInferSegmentsB200 returned nullptr")
        return buf.value.decode("utf-8")

    def model_forward(self, segments_f32, hw_emb=None):
        """funasr::Model::Forward(float**, int*, ..., hw_emb, ..., batch_in) -> list of result strings."""
        segs = [np.ascontiguousarray(s, dtype=np.float32) for s in segments_f32]
        n = len(segs)
        ptrs = (c_f32p * n)(*[s.ctypes.data_as(c_f32p) for s in segs])
        lens = np.asarray([len(s) for s in segs], np.int32)
        buf = C.create_string_buffer(1 << 22)
        if hw_emb is None:
            r = host_lib().b200pf_host_model_forward(self.h, ptrs, _p(lens, c_i32p), n, buf, len(buf))
        else:
            hw = np.ascontiguousarray(hw_emb, dtype=np.float32)
            r = host_lib().b200pf_host_model_forward_hw(self.h, ptrs, _p(lens, c_i32p), n, _p(hw), hw.shape[0],
                                                        hw.shape[1] if hw.ndim == 2 else 0, buf, len(buf))
        if r < 0:
            raise B200PFError("Model::Forward failed")
        return buf.value.decode("utf-8").split("\n")

    def model_forward_timed(self, segments_f32, iters=3):
        """Mean milliseconds per Model::Forward(float**, int*) call, timed inside the host library (no ctypes marshalling in the
        timed region), and the number of non-empty strings of the last call."""
        segs = [np.ascontiguousarray(s, dtype=np.float32) for s in segments_f32]
        n = len(segs)
        ptrs = (c_f32p * n)(*[s.ctypes.data_as(c_f32p) for s in segs])
        lens = np.asarray([len(s) for s in segs], np.int32)
        ms = C.c_double()
        H = host_lib()
        H.b200pf_host_model_forward_timed.argtypes = [C.c_void_p, C.POINTER(c_f32p), c_i32p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        r = H.b200pf_host_model_forward_timed(self.h, ptrs, _p(lens, c_i32p), n, int(iters), C.byref(ms))
        if r < 0:
            raise B200PFError("Model::Forward failed")
        return ms.value, r

    def segments_per_device(self):
        out = (C.c_longlong * 16)()
        n = host_lib().b200pf_host_segments_per_device(self.h, out, 16)
        return [int(out[i]) for i in range(n)]

    def init_seg_dict(self, path):
        host_lib().b200pf_host_init_seg_dict(self.h, path.encode())

    def compile_hotwords(self, hotwords, dim=512, cap_rows=4096):
        """CompileHotwordEmbedding(handle, hotwords) -> float32 [n_rows, dim]."""
        out = np.zeros((cap_rows, dim), np.float32)
        n = host_lib().b200pf_host_compile_hotwords(self.h, hotwords.encode("utf-8"), _p(out), cap_rows, dim)
        if n < 0:
            raise B200PFError("CompileHotwordEmbedding failed")
        return out[:n].copy()

    def infer_buffer_hw(self, pcm16, hw_emb, vad_max_len=60000):
        """FunOfflineInferBuffer with a hotword matrix -> (text, stamp)."""
        pcm16 = np.ascontiguousarray(pcm16, dtype="<i2")
        hw = np.ascontiguousarray(hw_emb, dtype=np.float32)
        t = C.create_string_buffer(1 << 20)
        st = C.create_string_buffer(1 << 20)
        n = host_lib().b200pf_host_offline_infer_buffer_hw(self.h, C.c_void_p(pcm16.ctypes.data), pcm16.nbytes, vad_max_len, _p(hw),
                                                           hw.shape[0], hw.shape[1], t, len(t), st, len(st))
        if n < 0:
            raise B200PFError("FunOfflineInferBuffer returned nullptr")
        return t.value.decode("utf-8"), st.value.decode("utf-8")


class TpassStream:
    """FunTpassInit + one FunTpassOnlineInit connection per `connect()` (the 2-pass server's handles); infer() = one
    FunTpassInferBuffer call -> dict(msg, tpass_msg, stamp, stamp_sents)."""

    def __init__(self, model_dir, vad_dir, punc_dir=None, device=0, options=None):
        kv = {"model-dir": model_dir, "vad-dir": vad_dir, "device": str(device)}
        if punc_dir:
            kv["punc-dir"] = punc_dir
        kv.update(options or {})
        keys = (C.c_char_p * len(kv))(*[k.encode() for k in kv])
        vals = (C.c_char_p * len(kv))(*[str(v).encode() for v in kv.values()])
        self.h = host_lib().b200pf_host_tpass_init_kv(keys, vals, len(kv))
        if not self.h:
            raise B200PFError("FunTpassInit failed")
        self._conns = []

    def connect(self):
        c = host_lib().b200pf_host_tpass_online_init(self.h)
        if not c:
            raise B200PFError("FunTpassOnlineInit failed")
        self._conns.append(c)
        return dict(h=c, cache=C.create_string_buffer(1 << 16))

    def infer(self, conn, pcm16, finished, mode=2, vad_tail_sil=800, vad_max_len=60000, cap=1 << 20):
        pcm16 = np.ascontiguousarray(pcm16, dtype="<i2")
        msg, tp, st, ss = (C.create_string_buffer(cap) for _ in range(4))
        n = host_lib().b200pf_host_tpass_infer(self.h, conn["h"], C.c_void_p(pcm16.ctypes.data), pcm16.nbytes, int(bool(finished)), mode, vad_tail_sil,
                                               vad_max_len, conn["cache"], len(conn["cache"]), msg, cap, tp, cap, st, cap, ss, cap)
        if n < 0:
            raise B200PFError("FunTpassInferBuffer returned nullptr")
        return dict(msg=msg.value.decode("utf-8", "replace"), tpass_msg=tp.value.decode("utf-8", "replace"), stamp=st.value.decode("utf-8"),
                    stamp_sents=ss.value.decode("utf-8", "replace"))

    def close(self):
        if self.h:
            for c in self._conns:
                host_lib().b200pf_host_tpass_online_uninit(c)
            self._conns = []
            host_lib().b200pf_host_tpass_uninit(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def host_expand_posteriors(topk_logprob, topk_ids, vocab):
    """pf::host::ExpandPrunedPosteriors: [rows, k] pruned posteriors -> dense [rows, vocab] log-softmax rows."""
    lp = np.ascontiguousarray(topk_logprob, dtype=np.float32)
    ids = np.ascontiguousarray(topk_ids, dtype=np.int32)
    rows, k = lp.shape
    out = np.zeros((rows, vocab), np.float32)
    host_lib().b200pf_host_expand_posteriors(_p(lp), _p(ids, c_i32p), rows, k, vocab, _p(out))
    return out


def host_vad_segments_streaming(sil_prob, chunk_lens, max_end_sil=800, max_seg_ms=15000, thres=0.8):
    """pf::host::StreamingVad fed in chunks of chunk_lens frames (the last chunk final) -> [(start_ms, end_ms)]."""
    p = np.ascontiguousarray(sil_prob, dtype=np.float32)
    cl = np.ascontiguousarray(chunk_lens, dtype=np.int32)
    out = np.zeros(2 * max(16, len(p) // 10 + 16), np.int32)
    n = host_lib().b200pf_host_vad_segments_streaming(_p(p), len(p), _p(cl, c_i32p), len(cl), max_end_sil, max_seg_ms, C.c_float(thres), _p(out, c_i32p),
                                                      len(out) // 2)
    return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(n)]


class MicroBatcher:
    """funasr_b200::MicroBatcher through its C hooks: merges concurrent batch-1 Forward calls (the 2-pass offline leg,
    funasrruntime.cpp:570-586) into batched forwards.  offline=None builds the host-only mock model (tests)."""

    def __init__(self, offline=None, max_wait_us=5000, max_batch=256, max_rows=32768, mock_latency_us=0):
        self._offline = offline
        if offline is None:
            self.h = host_lib().b200pf_host_mb_create_mock(max_wait_us, max_batch, max_rows, mock_latency_us)
        else:
            self.h = host_lib().b200pf_host_mb_create(offline.h, max_wait_us, max_batch, max_rows)
        if not self.h:
            raise B200PFError("MicroBatcher creation failed")

    def close(self):
        if self.h:
            host_lib().b200pf_host_mb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, pcm_f32, hw_emb=None):
        """Blocking batch-1 call of one connection; returns the result string."""
        x = np.ascontiguousarray(pcm_f32, dtype=np.float32)
        buf = C.create_string_buffer(1 << 16)
        if hw_emb is None:
            host_lib().b200pf_host_mb_forward(self.h, _p(x), len(x), None, 0, 0, buf, len(buf))
        else:
            hw = np.ascontiguousarray(hw_emb, dtype=np.float32)
            host_lib().b200pf_host_mb_forward(self.h, _p(x), len(x), _p(hw), hw.shape[0], hw.shape[1], buf, len(buf))
        return buf.value.decode("utf-8")

    def stats(self):
        out = (C.c_double * 7)()
        host_lib().b200pf_host_mb_stats(self.h, out)
        keys = ("segments", "batches", "closed_by_deadline", "closed_by_size", "max_batch_seen", "mean_wait_us", "max_wait_us")
        return dict(zip(keys, [float(v) for v in out]))


def funasr_infer(model_dir, wav_path=None, pcm16=None, device=0, max_rows=0):
    """FunASRInit -> FunASRInfer / FunASRInferBuffer -> FunASRGetResult -> FunASRUninit (the plain-model API)."""
    buf = C.create_string_buffer(1 << 20)
    if wav_path is not None:
        n = host_lib().b200pf_host_funasr_infer(model_dir.encode(), device, max_rows, wav_path.encode(), None, 0, buf, len(buf))
    else:
        a = np.ascontiguousarray(pcm16, dtype="<i2")
        n = host_lib().b200pf_host_funasr_infer(model_dir.encode(), device, max_rows, None, C.c_void_p(a.ctypes.data), a.nbytes, buf, len(buf))
    if n < 0:
        raise B200PFError("FunASRInit / FunASRInfer failed (%d)" % n)
    return buf.value.decode("utf-8")


def host_vad_segments(sil_prob, max_end_sil=800, max_seg_ms=15000, thres=0.8):
    """pf::host::SegmentVad -> int32 [n, 2] of [start_ms, end_ms]."""
    p = np.ascontiguousarray(sil_prob, dtype=np.float32)
    out = np.zeros((len(p) + 4, 2), np.int32)
    n = host_lib().b200pf_host_vad_segments(_p(p), len(p), int(max_end_sil), int(max_seg_ms), C.c_float(thres), _p(out, c_i32p), len(out))
    return out[:n].copy()


def host_pack_hotwords(tokens, hotwords, seg_dict_path=None, cap=4200):
    """funasr_b200::PackHotwords -> (ids int32 [n, 10], lengths int32 [n]); blank row last."""
    arr = (C.c_char_p * len(tokens))(*[t.encode("utf-8") for t in tokens])
    ids = np.zeros((cap, 10), np.int32)
    lens = np.zeros(cap, np.int32)
    n = host_lib().b200pf_host_pack_hotwords(arr, len(tokens), seg_dict_path.encode() if seg_dict_path else None,
                                             hotwords.encode("utf-8"), _p(ids, c_i32p), _p(lens, c_i32p), cap)
    if n < 0:
        raise B200PFError("too many hotwords")
    return ids[:n].copy(), lens[:n].copy()


def host_sentence_stamps(text, stamp):
    """pf::host::SentenceStamps (TimestampSentence)."""
    buf = C.create_string_buffer(64 * (len(text.encode("utf-8")) + len(stamp)) + 4096)
    n = host_lib().b200pf_host_sentence_stamps(text.encode("utf-8"), stamp.encode("utf-8"), buf, len(buf))
    assert n >= 0
    return buf.value.decode("utf-8", "replace")


class HostPuncTokenizer:
    """funasr_b200::PuncTokenizer + the AddPunc walk with a scripted network (CPU only)."""

    def __init__(self, tokens, punc_list):
        self.tokens = (C.c_char_p * len(tokens))(*[t.encode("utf-8") for t in tokens])
        self.n = len(tokens)
        self.punc = (C.c_char_p * len(punc_list))(*[t.encode("utf-8") for t in punc_list])
        self.n_punc = len(punc_list)

    def tokenize(self, text):
        raw = text.encode("utf-8")
        ids = np.zeros(len(raw) + 4, np.int32)
        n = host_lib().b200pf_host_punc_tokenize(self.tokens, self.n, raw, _p(ids, c_i32p), len(ids))
        assert n >= 0
        return ids[:n].tolist()

    def add_punc_online_scripted(self, text, cache, seed, every):
        """The realtime walk (AddPuncOnlineWith) with the scripted network seeded by seed + 13 * vad_pos; cache updated in place."""
        raw, cin = text.encode("utf-8"), _cache_join(cache)
        cap = 16 * (len(raw) + len(cin)) + 4096
        buf, cbuf = C.create_string_buffer(cap), C.create_string_buffer(cap)
        n = host_lib().b200pf_host_punc_online_add_scripted(self.tokens, self.n, self.punc, self.n_punc, raw, cin, seed, every, buf, cap, cbuf, cap)
        assert n >= 0
        cache[:] = _cache_split(cbuf.value)
        return buf.value.decode("utf-8", "replace")

    def add_punc_scripted(self, text, lang, seed, every):
        raw = text.encode("utf-8")
        buf = C.create_string_buffer(16 * len(raw) + 4096)
        n = host_lib().b200pf_host_punc_add_scripted(self.tokens, self.n, self.punc, self.n_punc, raw, lang.encode(), seed, every, buf, len(buf))
        assert n >= 0
        return buf.value.decode("utf-8", "replace")


def _cache_join(cache):
    return b"".join(w + b"\x01" for w in cache)


def _cache_split(raw):
    return raw.split(b"\x01")[:-1]


class HostPuncOnline:
    """funasr_b200::CTTransformerOnlineB200 (the realtime punctuation model): add_punc(text, cache) with cache a list of bytes,
    updated in place."""

    def __init__(self, punc_dir, device=0, max_tokens=0):
        self.h = host_lib().b200pf_host_punc_online_create(punc_dir.encode(), device, max_tokens)
        if not self.h:
            raise B200PFError("CTTransformerOnline init failed: " + lib().b200pf_last_error().decode("utf-8", "replace"))

    def close(self):
        if self.h:
            host_lib().b200pf_host_punc_online_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_punc(self, text, cache):
        raw, cin = text.encode("utf-8"), _cache_join(cache)
        cap = 16 * (len(raw) + len(cin)) + 4096
        buf, cbuf = C.create_string_buffer(cap), C.create_string_buffer(cap)
        n = host_lib().b200pf_host_punc_online_add(self.h, raw, cin, buf, cap, cbuf, cap)
        if n < 0:
            raise B200PFError("AddPunc (online) failed")
        cache[:] = _cache_split(cbuf.value)
        return buf.value.decode("utf-8", "replace")


class HostPunc:
    """CTTransformerInit / CTTransformerInfer / CTTransformerUninit through the host shim (funasr_b200::CTTransformerB200)."""

    def __init__(self, punc_dir, device=0, max_tokens=0):
        self.h = host_lib().b200pf_host_punc_create(punc_dir.encode(), device, max_tokens)
        if not self.h:
            raise B200PFError("CTTransformerInit failed: " + lib().b200pf_last_error().decode("utf-8", "replace"))

    def close(self):
        if self.h:
            host_lib().b200pf_host_punc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def rounds(self):
        return int(host_lib().b200pf_host_punc_rounds(self.h))

    def add_punc(self, text, lang="zh-cn"):
        raw = text.encode("utf-8")
        buf = C.create_string_buffer(16 * len(raw) + 4096)
        n = host_lib().b200pf_host_punc_add(self.h, raw, lang.encode(), buf, len(buf))
        if n < 0:
            raise B200PFError("AddPunc failed")
        return buf.value.decode("utf-8", "replace")

    def add_punc_batch(self, texts, lang="zh-cn"):
        """-> (list of punctuated texts, engine calls made)"""
        raws = [t.encode("utf-8") for t in texts]
        arr = (C.c_char_p * len(raws))(*raws)
        cap = 16 * sum(len(r) for r in raws) + 4096 * (len(raws) + 1)
        buf = C.create_string_buffer(cap)
        rounds = np.zeros(1, np.int32)
        n = host_lib().b200pf_host_punc_add_batch(self.h, arr, len(raws), lang.encode(), buf, cap, _p(rounds, c_i32p))
        if n < 0:
            raise B200PFError("AddPuncBatch failed")
        parts = buf.raw[:n].split(b"\0")[:len(raws)]
        return [p.decode("utf-8", "replace") for p in parts], int(rounds[0])


def host_partition(lens, n_dev):
    """MultiGpuParaformer's LPT assignment of segments (sample counts) to n_dev queues."""
    l = np.ascontiguousarray(lens, dtype=np.int32)
    a = np.zeros(len(l), np.int32)
    host_lib().b200pf_host_partition(_p(l, c_i32p), len(l), n_dev, _p(a, c_i32p))
    return a


def _check(rc):
    if rc != 0:
        raise B200PFError("b200pf error %d: %s" % (rc, lib().b200pf_last_error().decode("utf-8", "replace")))


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a, typ=c_f32p):
    return None if a is None else a.ctypes.data_as(typ)


def model_dir_probe(model_dir):
    cfg = Config()
    nt, nw = C.c_int32(), C.c_int32()
    _check(lib().b200pf_model_dir_probe(model_dir.encode(), C.byref(cfg), C.byref(nt), C.byref(nw)))
    return cfg, nt.value, nw.value


def device_count():
    return lib().b200pf_device_count()


class Engine:
    """One GPU: resident 16-bit weights + workspace (replaces Paraformer::InitAsr, paraformer.cpp:21-53).
    prec: None = the library default (fp16 operands; env B200PF_PREC overrides), "bf16" / "fp16" or PREC_* to force one."""

    def __init__(self, model_dir, device=0, max_rows=0, max_segments=0, prec=None):
        self.h = C.c_void_p()
        p = -1 if prec is None else ({"bf16": PREC_BF16, "fp16": PREC_FP16}[prec] if isinstance(prec, str) else int(prec))
        _check(lib().b200pf_engine_create_prec(model_dir.encode(), device, max_rows, max_segments, p, C.byref(self.h)))
        self.cfg = Config()
        _check(lib().b200pf_engine_config(self.h, C.byref(self.cfg)))

    def close(self):
        if self.h:
            for ref in list(getattr(self, "_batches", [])):
                b = ref()
                if b is not None:
                    b.close()          # a batch must not outlive its engine
            lib().b200pf_engine_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        _check(lib().b200pf_engine_set_option(self.h, key.encode(), int(value)))
        if key == "logprob_topk":
            self._topk = int(value)

    def graph_stats(self):
        """(graphs captured, graph replays, graphs cached) of the small-batch CUDA-graph path."""
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_int()
        lib().b200pf_engine_graph_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
        _check(lib().b200pf_engine_graph_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(captures=a.value, replays=b.value, cached=c.value)

    def profile_read(self, reset=True):
        names = (C.c_char_p * 16)()
        ms = (C.c_double * 16)()
        work = (C.c_double * 16)()
        n = (C.c_longlong * 16)()
        _check(lib().b200pf_engine_profile_read(self.h, int(reset), names, ms, work, n))
        return {names[i].decode(): dict(ms=ms[i], work=work[i], launches=int(n[i])) for i in range(16) if names[i]}

    @property
    def stream(self):
        return lib().b200pf_engine_stream(self.h)

    def tokens(self):
        n = lib().b200pf_engine_vocab_size(self.h)
        return [lib().b200pf_engine_token(self.h, i).decode("utf-8") for i in range(n)]

    @property
    def lang(self):
        return lib().b200pf_engine_lang(self.h).decode()

    def hotword_embed(self, ids, lengths):
        """Embedding + LSTM of the hotword compiler: ids int32 [n, max_len], lengths [n] -> float32 [n, d_model]."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        out = np.zeros((ids.shape[0], self.cfg.d_model), np.float32)
        _check(lib().b200pf_engine_hotword_embed(self.h, _p(ids, c_i32p), _p(lengths, c_i32p), ids.shape[0], ids.shape[1], _p(out)))
        return out

    def frontend(self, pcm16):
        pcm16 = np.ascontiguousarray(pcm16, dtype=np.int16)
        nfb = lib().b200pf_num_fbank_frames(len(pcm16))
        T = lib().b200pf_num_lfr_frames(len(pcm16))
        fb = np.zeros((nfb, 80), np.float32)
        feats = np.zeros((T, 560), np.float32)
        if nfb:
            _check(lib().b200pf_op_frontend(self.h, _p(pcm16, c_i16p), len(pcm16), _p(fb), _p(feats)))
        return fb, feats


class VadEngine:
    """FSMN-VAD scores on the GPU (replaces FsmnVad::Forward's onnxruntime session, fsmn-vad.cpp:72-135)."""

    def __init__(self, vad_dir, device=0, max_frames=0):
        self.h = C.c_void_p()
        lib().b200pf_vad_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        lib().b200pf_vad_destroy.argtypes = [C.c_void_p]
        lib().b200pf_vad_destroy.restype = None
        lib().b200pf_vad_scores_s16.argtypes = [C.c_void_p, C.c_void_p, c_i64p, C.c_int, c_f32p, C.c_int64, c_i32p, c_f32p, c_f32p]
        _check(lib().b200pf_vad_create(vad_dir.encode(), device, max_frames, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().b200pf_vad_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def scores(self, pcm16, offsets, all_probs=False, feats=False):
        """-> (sil_prob [F], frame_off [n+1], probs [F,248] or None, feats [F,400] or None)"""
        pcm16 = np.ascontiguousarray(pcm16, dtype=np.int16)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        cap = int(sum(lib().b200pf_num_fbank_frames(int(offsets[i + 1] - offsets[i])) for i in range(n)))
        p0 = np.zeros(max(cap, 1), np.float32)
        fo = np.zeros(n + 1, np.int32)
        pr = np.zeros((max(cap, 1), 248), np.float32) if all_probs else None
        ft = np.zeros((max(cap, 1), 400), np.float32) if feats else None
        _check(lib().b200pf_vad_scores_s16(self.h, C.c_void_p(pcm16.ctypes.data), _p(offsets, c_i64p), n, _p(p0), cap, _p(fo, c_i32p),
                                           _p(pr), _p(ft)))
        return p0[:cap], fo, (pr[:cap] if pr is not None else None), (ft[:cap] if ft is not None else None)


class PuncEngine:
    """CT-Transformer punctuation network on the GPU (replaces CTTransformer::Infer's onnxruntime session, ct-transformer.cpp:164-203)."""

    def __init__(self, punc_dir, device=0, max_tokens=0):
        self.h = C.c_void_p()
        L = lib()
        L.b200pf_punc_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.b200pf_punc_destroy.argtypes = [C.c_void_p]
        L.b200pf_punc_destroy.restype = None
        L.b200pf_punc_info.argtypes = [C.c_void_p, c_i32p, c_i32p, c_i32p, c_i32p]
        L.b200pf_punc_infer.argtypes = [C.c_void_p, c_i32p, c_i32p, C.c_int, c_i32p, c_f32p]
        L.b200pf_punc_infer_vad.argtypes = [C.c_void_p, c_i32p, c_i32p, c_i32p, C.c_int, c_i32p, c_f32p]
        L.b200pf_punc_launches.argtypes = [C.c_void_p]
        L.b200pf_punc_launches.restype = C.c_longlong
        _check(L.b200pf_punc_create(punc_dir.encode(), device, max_tokens, C.byref(self.h)))
        info = np.zeros(4, np.int32)
        _check(L.b200pf_punc_info(self.h, _p(info[0:1], c_i32p), _p(info[1:2], c_i32p), _p(info[2:3], c_i32p), _p(info[3:4], c_i32p)))
        self.vocab, self.n_punc, self.d_model, self.max_tokens = (int(v) for v in info)

    def close(self):
        if self.h:
            lib().b200pf_punc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(lib().b200pf_punc_launches(self.h))

    def infer(self, ids, offsets, logits=False, vad_pos=None):
        """-> (punc ids int32 [T], logits [T, n_punc] or None) for sequences ids[offsets[i]:offsets[i+1]]; vad_pos [n_seq]: the
        realtime model's VadMask position per sequence (b200pf_punc_infer_vad)."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        out = np.zeros(max(1, len(ids)), np.int32)
        lg = np.zeros((max(1, len(ids)), self.n_punc), np.float32) if logits else None
        if vad_pos is not None:
            vp = np.ascontiguousarray(vad_pos, dtype=np.int32)
            assert len(vp) == len(offsets) - 1
            _check(lib().b200pf_punc_infer_vad(self.h, _p(ids, c_i32p), _p(offsets, c_i32p), _p(vp, c_i32p), len(offsets) - 1, _p(out, c_i32p), _p(lg)))
            return out[:len(ids)], (lg[:len(ids)] if lg is not None else None)
        _check(lib().b200pf_punc_infer(self.h, _p(ids, c_i32p), _p(offsets, c_i32p), len(offsets) - 1, _p(out, c_i32p), _p(lg)))
        return out[:len(ids)], (lg[:len(ids)] if lg is not None else None)


class Batch:
    """One batch of segments (device PCM + packed layout + results)."""

    def __init__(self, engine, max_samples):
        self.engine = engine
        self.h = C.c_void_p()
        _check(lib().b200pf_batch_create(engine.h, int(max_samples), C.byref(self.h)))
        import weakref
        if not hasattr(engine, "_batches"):
            engine._batches = []
        engine._batches.append(weakref.ref(self))
        self._keep = None
        self.n_seg = 0

    def close(self):
        if self.h:
            lib().b200pf_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stage_s16(self, pcm16, offsets, stream=None):
        """pcm16: int16 numpy array or an integer host address; offsets: int64 [n_seg+1]."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        if isinstance(pcm16, np.ndarray):
            pcm16 = np.ascontiguousarray(pcm16, dtype=np.int16)
            ptr = pcm16.ctypes.data
        else:
            ptr = int(pcm16)
        self._keep = (pcm16, offsets)
        self.n_seg = len(offsets) - 1
        _check(lib().b200pf_batch_stage_s16(self.h, C.c_void_p(ptr), _p(offsets, c_i64p), self.n_seg, C.c_void_p(stream or 0)))

    def stage_f32(self, segments, stream=None):
        segs = [np.ascontiguousarray(s, dtype=np.float32) for s in segments]
        n = len(segs)
        ptrs = (c_f32p * n)(*[s.ctypes.data_as(c_f32p) for s in segs])
        lens = np.asarray([len(s) for s in segs], np.int32)
        self._keep = (segs, ptrs, lens)
        self.n_seg = n
        _check(lib().b200pf_batch_stage_f32(self.h, ptrs, _p(lens, c_i32p), n, C.c_void_p(stream or 0)))

    def set_hotwords(self, hw_emb):
        """hw_emb float32 [n_hw, d_model] (what Forward receives as hw_emb); call before stage_*."""
        hw = np.ascontiguousarray(hw_emb, dtype=np.float32)
        _check(lib().b200pf_batch_set_hotwords(self.h, _p(hw), hw.shape[0], hw.shape[1] if hw.ndim == 2 else 0))

    def run(self, stream=None):
        _check(lib().b200pf_batch_run(self.h, C.c_void_p(stream or 0)))

    def collect(self, cap_tokens=None, stream=None):
        n = self.n_seg
        cap = int(cap_tokens or self.engine.cfg.max_rows)
        out = dict(token_counts=np.zeros(n, np.int32), token_offsets=np.zeros(n + 1, np.int32),
                   lfr_frames=np.zeros(n, np.int32), token_ids=np.zeros(cap, np.int32), fire_frames=np.zeros(cap, np.int32))
        out["us_offsets"] = np.zeros(n + 1, np.int32)
        ts = bool(self.engine.cfg.timestamp)
        cap_us = 3 * int(self.engine.cfg.max_rows) if ts else 0
        if ts:
            out["us_alphas"] = np.zeros(cap_us, np.float32)
            out["us_peaks"] = np.zeros(cap_us, np.float32)
        r = Result(_p(out["token_counts"], c_i32p), _p(out["token_offsets"], c_i32p), _p(out["lfr_frames"], c_i32p),
                   _p(out["token_ids"], c_i32p), _p(out["fire_frames"], c_i32p), cap, 0,
                   _p(out.get("us_alphas")), _p(out.get("us_peaks")), _p(out["us_offsets"], c_i32p), cap_us, None, None, None, 0)
        k = int(getattr(self.engine, "_topk", 0))
        if k > 0:
            out["token_lse"] = np.zeros(cap, np.float32)
            out["topk_logprob"] = np.zeros((cap, k), np.float32)
            out["topk_ids"] = np.zeros((cap, k), np.int32)
            r.token_lse, r.topk_logprob, r.topk_ids = _p(out["token_lse"]), _p(out["topk_logprob"]), _p(out["topk_ids"], c_i32p)
        _check(lib().b200pf_batch_collect(self.h, C.byref(r), C.c_void_p(stream or 0)))
        if k > 0:
            for key in ("token_lse", "topk_logprob", "topk_ids"):
                out[key] = out[key][:int(r.n_tokens)]
        if ts:
            nu = int(out["us_offsets"][n])
            out["us_alphas"] = out["us_alphas"][:nu]
            out["us_peaks"] = out["us_peaks"][:nu]
        nt = int(r.n_tokens)
        out["token_ids"] = out["token_ids"][:nt]
        out["fire_frames"] = out["fire_frames"][:nt]
        out["n_tokens"] = nt
        return out

    def forward_s16(self, pcm16, offsets):
        self.stage_s16(pcm16, offsets)
        self.run()
        return self.collect()

    def forward_f32(self, segments):
        self.stage_f32(segments)
        self.run()
        return self.collect()

    @property
    def launches(self):
        return int(lib().b200pf_batch_launches(self.h))

    @property
    def flops(self):
        return float(lib().b200pf_batch_flops(self.h))

    def tap(self, name, seg, cap=None):
        cfg = self.engine.cfg
        cap = int(cap or (2100 * max(cfg.vocab, 560)))
        buf = np.zeros(cap, np.float32)
        shape = (C.c_int64 * 2)()
        _check(lib().b200pf_batch_tap(self.h, name.encode(), int(seg), _p(buf), cap, shape))
        r, c = int(shape[0]), int(shape[1])
        a = buf[: r * c].copy()
        return a.reshape(r) if c == 1 else a.reshape(r, c)


# ---- single-operator wrappers (parity tests) ---------------------------------------------------------
def op_set_precision(prec):
    """16-bit operand format of the op_* entry points of this process: "bf16" / "fp16" (the default, like the engine)."""
    _check(lib().b200pf_op_set_precision({"bf16": PREC_BF16, "fp16": PREC_FP16}[prec] if isinstance(prec, str) else int(prec)))


def op_gemm(A, W, bias=None, add=None, res=None, relu=0, out_bf16=False, argmax=False, device=0, general=False):
    A, W = _f32(A), _f32(W)
    M, K = A.shape
    N = W.shape[0]
    bias, add, res = _f32(bias), _f32(add), _f32(res)
    out = np.zeros((M, N), np.float32)
    am = np.zeros(M, np.int32) if argmax else None
    _check(lib().b200pf_op_gemm(device, _p(A), _p(W), _p(bias), _p(add), _p(res), M, N, K, int(relu), 2 if general else int(out_bf16),
                                _p(out), _p(am, c_i32p)))
    return (out, am) if argmax else out


def op_lstm(x, seq_off, seq_len, w_ih, w_hh, b_ih, b_hh, bf16_out=False, device=0):
    """x [rows,512]; weights [n_dir*2048,512] (forward then reverse), biases [n_dir*2048] -> [rows, 512*n_dir]."""
    x, w_ih, w_hh, b_ih, b_hh = _f32(x), _f32(w_ih), _f32(w_hh), _f32(b_ih), _f32(b_hh)
    n_dir = w_ih.shape[0] // 2048
    so = np.ascontiguousarray(seq_off, dtype=np.int32)
    sl = np.ascontiguousarray(seq_len, dtype=np.int32)
    out = np.zeros((x.shape[0], 512 * n_dir), np.float32)
    _check(lib().b200pf_op_lstm(device, _p(x), x.shape[0], _p(so, c_i32p), _p(sl, c_i32p), len(so), n_dir, _p(w_ih), _p(w_hh),
                                _p(b_ih), _p(b_hh), int(bf16_out), _p(out)))
    return out


def op_logprob_topk(logits, k, device=0):
    x = _f32(logits)
    rows, V = x.shape
    lse = np.zeros(rows, np.float32)
    lp = np.zeros((rows, k), np.float32)
    ids = np.zeros((rows, k), np.int32)
    _check(lib().b200pf_op_logprob_topk(device, _p(x), rows, V, k, _p(lse), _p(lp), _p(ids, c_i32p)))
    return lse, lp, ids


def op_lstm_bench(n_seq, length, n_dir=1, iters=3, device=0):
    ms = C.c_float()
    mc = C.c_int()
    _check(lib().b200pf_op_lstm_bench(device, n_seq, length, n_dir, iters, C.byref(ms), C.byref(mc)))
    return ms.value, mc.value


def op_us_peaks(alpha2, seq_off, seq_len, n_tok, threshold, device=0):
    a = _f32(alpha2)
    so = np.ascontiguousarray(seq_off, dtype=np.int32)
    sl = np.ascontiguousarray(seq_len, dtype=np.int32)
    nt = np.ascontiguousarray(n_tok, dtype=np.int32)
    ua, up = np.zeros_like(a), np.zeros_like(a)
    _check(lib().b200pf_op_us_peaks(device, _p(a), _p(so, c_i32p), _p(sl, c_i32p), _p(nt, c_i32p), len(so), len(a),
                                    C.c_float(threshold), _p(ua), _p(up)))
    return ua, up


def op_gemm_bench(M, N, K, mode=0, iters=20, device=0):
    ms = C.c_float()
    _check(lib().b200pf_op_gemm_bench(device, M, N, K, mode, iters, C.byref(ms)))
    return ms.value


def op_conv3(X, Wr, bias, device=0):
    X, Wr, bias = _f32(X), _f32(Wr), _f32(bias)
    M, Cc = X.shape
    out = np.zeros((M, Cc), np.float32)
    _check(lib().b200pf_op_conv3(device, _p(X), _p(Wr), _p(bias), M, Cc, _p(out)))
    return out


def op_layernorm(x, gamma, beta, eps=1e-12, in_bf16=False, device=0):
    x, gamma, beta = _f32(x), _f32(gamma), _f32(beta)
    rows, D = x.shape
    o32 = np.zeros((rows, D), np.float32)
    o16 = np.zeros((rows, D), np.float32)
    _check(lib().b200pf_op_layernorm(device, _p(x), rows, D, _p(gamma), _p(beta), eps, int(in_bf16), _p(o32), _p(o16)))
    return o32, o16


def op_attention(q, k, v, q_off, q_len, kv_off, kv_len, n_heads=4, impl=0, device=0):
    q, k, v = _f32(q), _f32(k), _f32(v)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    q_off, q_len, kv_off, kv_len = i32(q_off), i32(q_len), i32(kv_off), i32(kv_len)
    out = np.zeros_like(q)
    _check(lib().b200pf_op_attention(device, _p(q), _p(k), _p(v), _p(q_off, c_i32p), _p(q_len, c_i32p), _p(kv_off, c_i32p),
                                     _p(kv_len, c_i32p), len(q_len), n_heads, q.shape[0], k.shape[0], impl, _p(out)))
    return out


def op_attention_bench(seg_T, n_heads=4, cross=False, impl=0, iters=20, device=0):
    """(ms per launch, algorithmic FLOPs per launch) of the attention kernel on engine-shaped random operands."""
    t = np.ascontiguousarray(seg_T, dtype=np.int32)
    ms, fl = C.c_float(), C.c_double()
    _check(lib().b200pf_op_attention_bench(device, _p(t, c_i32p), len(t), n_heads, int(bool(cross)), impl, iters, C.byref(ms), C.byref(fl)))
    return ms.value, fl.value


def op_fsmn(x, w, seg_off, device=0):
    x, w = _f32(x), _f32(w).reshape(512, 11)
    seg_off = np.ascontiguousarray(seg_off, dtype=np.int32)
    out = np.zeros_like(x)
    _check(lib().b200pf_op_fsmn(device, _p(x), _p(w), _p(seg_off, c_i32p), len(seg_off) - 1, _p(out)))
    return out


def op_cif(alphas, hidden, seg_off, threshold=1.0, device=0):
    alphas, hidden = _f32(alphas), _f32(hidden)
    seg_off = np.ascontiguousarray(seg_off, dtype=np.int32)
    n = len(seg_off) - 1
    rows = len(alphas)
    n_tok = np.zeros(n, np.int32)
    fires = np.zeros(rows, np.float32)
    emb = np.zeros((rows, 512), np.float32)
    ff = np.zeros(rows, np.int32)
    _check(lib().b200pf_op_cif(device, _p(alphas), _p(hidden), _p(seg_off, c_i32p), n, threshold, _p(n_tok, c_i32p),
                               _p(fires), _p(emb), _p(ff, c_i32p), rows))
    tot = int(n_tok.sum())
    return n_tok, fires, emb[:tot], ff[:tot]
