"""ctypes binding of ``libsnacb.so`` (include/snacb.h).  No fallback: if the CUDA library is
missing or cannot be loaded, importing callers get a RuntimeError."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libsnacb.so")

RAW_IDS, EXTRACT_SLICE, FP32, KEEP_TAPS, STREAM_FP32, BF16, UNFUSED = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40

_FP = C.POINTER(C.c_float)


class ResUnitWeights(C.Structure):
    _fields_ = [(n, _FP) for n in ("alpha1", "dw_w", "dw_b", "alpha2", "pw_w", "pw_b")]


class BlockWeights(C.Structure):
    _fields_ = [("alpha", _FP), ("convt_w", _FP), ("convt_b", _FP), ("noise_w", _FP), ("res", ResUnitWeights * 3)]


class Weights(C.Structure):
    _fields_ = [
        ("codebook", _FP * 3), ("out_proj_w", _FP * 3), ("out_proj_b", _FP * 3),
        ("stem_dw_w", _FP), ("stem_dw_b", _FP), ("stem_pw_w", _FP), ("stem_pw_b", _FP),
        ("block", BlockWeights * 4),
        ("tail_alpha", _FP), ("tail_w", _FP), ("tail_b", _FP),
    ]


class EncBlockWeights(C.Structure):
    _fields_ = [("alpha", _FP), ("conv_w", _FP), ("conv_b", _FP), ("res", ResUnitWeights * 3)]


class EncoderWeights(C.Structure):
    _fields_ = [
        ("conv0_w", _FP), ("conv0_b", _FP), ("block", EncBlockWeights * 4), ("final_w", _FP), ("final_b", _FP),
        ("in_proj_w", _FP * 3), ("in_proj_b", _FP * 3), ("codebook", _FP * 3), ("out_proj_w", _FP * 3), ("out_proj_b", _FP * 3),
    ]


EXPORTS = [
    "snacb_version", "snacb_create", "snacb_destroy", "snacb_last_error", "snacb_unpack", "snacb_decode", "snacb_decode_keyed", "snacb_decode_range",
    "snacb_decode_host", "snacb_decode_host_submit", "snacb_decode_host_wait", "snacb_samples_out", "snacb_set_group_bytes", "snacb_stats",
    "snacb_profile", "snacb_profile_report", "snacb_debug_tap_count", "snacb_debug_tap_info", "snacb_debug_tap_copy", "snacb_debug_chain_spans", "snacb_debug_chain_spans_ex", "snacb_debug_chain_plan", "snacb_debug_chain_ws_spans", "snacb_chain_modes",
    "snacb_batcher_create", "snacb_batcher_destroy", "snacb_batcher_push", "snacb_batcher_end",
    "snacb_batcher_flush", "snacb_batcher_pending", "snacb_batcher_forget", "snacb_batcher_flush_submit", "snacb_batcher_flush_wait", "snacb_batcher_take",
    "snacb_ingest_create", "snacb_ingest_destroy", "snacb_ingest_reset", "snacb_ingest_window_capacity",
    "snacb_ingest_step", "snacb_ingest_state", "snacb_base64_len", "snacb_pcm_to_base64", "snacb_pcm_to_wav",
    "snacb_session_create", "snacb_session_destroy", "snacb_session_bytes", "snacb_session_max_frames", "snacb_session_reset",
    "snacb_session_frames", "snacb_session_emitted", "snacb_session_next_emit", "snacb_session_step", "snacb_session_step_multi",
    "snacb_debug_session_frontier", "snacb_experiments_built",
    "snacb_encoder_create", "snacb_encoder_destroy", "snacb_encoder_last_error", "snacb_encoder_launches", "snacb_encode_frames",
    "snacb_encode", "snacb_pack_tokens",
    "snacb_streamer_create", "snacb_streamer_destroy", "snacb_streamer_push", "snacb_streamer_end", "snacb_streamer_active",
    "snacb_streamer_tick",
]

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m tts_inference_b200.build` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32p, i16p, u64 = C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64
    lib.snacb_version.restype = C.c_int
    lib.snacb_create.argtypes = [C.POINTER(vp), C.POINTER(Weights), C.c_int]
    lib.snacb_destroy.argtypes = [vp]
    lib.snacb_destroy.restype = None
    lib.snacb_last_error.argtypes = [vp]
    lib.snacb_last_error.restype = C.c_char_p
    lib.snacb_unpack.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int, i32p, i32p, i32p, vp]
    lib.snacb_decode.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp), u64, i16p, vp, vp]
    lib.snacb_decode_keyed.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp), u64, i32p, i16p, vp, vp]
    lib.snacb_decode_range.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp), u64, i32p, C.c_int, C.c_int,
                                       i16p, vp, vp]
    lib.snacb_decode_host.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int, C.c_int, u64, i16p]
    lib.snacb_decode_host_submit.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int, C.c_int, u64, i16p]
    lib.snacb_decode_host_wait.argtypes = [vp]
    lib.snacb_samples_out.argtypes = [C.c_int, C.c_int]
    lib.snacb_set_group_bytes.argtypes = [vp, C.c_size_t]
    lib.snacb_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    lib.snacb_profile.argtypes = [vp, C.c_int]
    lib.snacb_profile_report.argtypes = [vp, C.c_char_p, C.c_size_t]
    lib.snacb_debug_tap_count.argtypes = [vp]
    lib.snacb_debug_tap_info.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.snacb_debug_tap_copy.argtypes = [vp, C.c_int, vp, C.c_size_t]
    lib.snacb_debug_chain_spans.argtypes = [C.c_int, i16p, C.c_int]
    lib.snacb_debug_chain_spans_ex.argtypes = [C.c_int, C.c_int, C.c_int, i16p, C.c_int]
    lib.snacb_debug_chain_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    lib.snacb_debug_chain_ws_spans.argtypes = [C.c_int, i16p, C.c_int]
    lib.snacb_chain_modes.argtypes = [vp, i32p]
    lib.snacb_batcher_create.argtypes = [C.POINTER(vp), vp, C.c_int, C.c_int, C.c_int]
    lib.snacb_batcher_destroy.argtypes = [vp]
    lib.snacb_batcher_destroy.restype = None
    lib.snacb_batcher_push.argtypes = [vp, u64, vp, C.c_int]
    lib.snacb_batcher_end.argtypes = [vp, u64]
    lib.snacb_batcher_flush.argtypes = [vp, u64, C.c_int, vp, vp, vp, vp, C.c_size_t]
    lib.snacb_batcher_flush_submit.argtypes = [vp, u64, C.c_int, vp, vp, vp, vp, C.c_size_t]
    lib.snacb_batcher_flush_wait.argtypes = [vp]
    lib.snacb_batcher_forget.argtypes = [vp, u64]
    lib.snacb_batcher_take.argtypes = [vp, C.c_int, vp, i32p, i32p]
    lib.snacb_batcher_pending.argtypes = [vp]
    lib.snacb_ingest_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    lib.snacb_ingest_destroy.argtypes = [vp]
    lib.snacb_ingest_destroy.restype = None
    lib.snacb_ingest_reset.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.snacb_ingest_window_capacity.argtypes = [C.c_int, C.c_int]
    lib.snacb_ingest_step.argtypes = [vp, i32p, C.c_int, C.c_int, i32p, vp, i32p, i32p, C.c_int, i32p, i32p, i32p, i32p, vp]
    lib.snacb_ingest_state.argtypes = [vp, vp, vp, C.c_int]
    lib.snacb_base64_len.argtypes = [C.c_longlong]
    lib.snacb_base64_len.restype = C.c_longlong
    lib.snacb_pcm_to_base64.argtypes = [i16p, C.c_longlong, C.c_longlong, vp, vp]
    lib.snacb_pcm_to_wav.argtypes = [i16p, C.c_longlong, C.c_longlong, C.c_int, vp, vp]
    lib.snacb_session_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.snacb_session_destroy.argtypes = [vp]
    lib.snacb_session_destroy.restype = None
    lib.snacb_session_bytes.argtypes = [vp]
    lib.snacb_session_bytes.restype = C.c_int64
    lib.snacb_session_max_frames.argtypes = [vp]
    lib.snacb_session_reset.argtypes = [vp, C.c_int, C.c_int]
    lib.snacb_session_frames.argtypes = [vp, C.c_int]
    lib.snacb_session_frames.restype = C.c_int64
    lib.snacb_session_emitted.argtypes = [vp, C.c_int]
    lib.snacb_session_emitted.restype = C.c_int64
    lib.snacb_session_next_emit.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.snacb_session_step.argtypes = [vp, C.c_int, C.c_int, i32p, C.c_int, C.c_int, C.c_int, u64, i32p, i16p, C.c_int,
                                       C.POINTER(C.c_int), vp]
    lib.snacb_session_step_multi.argtypes = [vp, C.c_int, i32p, i32p, C.c_int, C.c_int, u64, i32p, i16p, C.c_int,
                                             C.POINTER(C.c_int), vp]
    lib.snacb_debug_session_frontier.argtypes = [C.c_int, C.c_int, i32p, C.c_int]
    lib.snacb_encoder_create.argtypes = [C.POINTER(vp), C.POINTER(EncoderWeights), C.c_int]
    lib.snacb_encoder_destroy.argtypes = [vp]
    lib.snacb_encoder_destroy.restype = None
    lib.snacb_encoder_last_error.argtypes = [vp]
    lib.snacb_encoder_last_error.restype = C.c_char_p
    lib.snacb_encoder_launches.argtypes = [vp]
    lib.snacb_encoder_launches.restype = C.c_uint64
    lib.snacb_encode_frames.argtypes = [C.c_int]
    lib.snacb_encode.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, i32p, i32p, i32p, vp, vp, vp]
    lib.snacb_pack_tokens.argtypes = [i32p, i32p, i32p, C.c_int, C.c_int, C.c_int, i32p, vp]
    lib.snacb_streamer_create.argtypes = [C.POINTER(vp), vp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.snacb_streamer_destroy.argtypes = [vp]
    lib.snacb_streamer_destroy.restype = None
    lib.snacb_streamer_push.argtypes = [vp, u64, vp, C.c_int]
    lib.snacb_streamer_end.argtypes = [vp, u64]
    lib.snacb_streamer_active.argtypes = [vp]
    lib.snacb_streamer_tick.argtypes = [vp, u64, C.c_int, vp, vp, vp, vp, C.c_size_t]
    _lib = lib
    return lib


def make_weights(folded: Dict[str, np.ndarray]):
    """``snacb_weights`` struct pointing into ``folded`` (returned alongside to keep it alive)."""
    def p(k):
        a = folded[k]
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], k
        return a.ctypes.data_as(_FP)
    w = Weights()
    for i in range(3):
        w.codebook[i] = p(f"codebook{i}")
        w.out_proj_w[i] = p(f"out_proj_w{i}")
        w.out_proj_b[i] = p(f"out_proj_b{i}")
    w.stem_dw_w, w.stem_dw_b = p("stem_dw_w"), p("stem_dw_b")
    w.stem_pw_w, w.stem_pw_b = p("stem_pw_w"), p("stem_pw_b")
    for bi in range(4):
        b = w.block[bi]
        b.alpha, b.convt_w, b.convt_b, b.noise_w = p(f"b{bi}.alpha"), p(f"b{bi}.convt_w"), p(f"b{bi}.convt_b"), p(f"b{bi}.noise_w")
        for ri in range(3):
            r = b.res[ri]
            for n in ("alpha1", "dw_w", "dw_b", "alpha2", "pw_w", "pw_b"):
                setattr(r, n, p(f"b{bi}.r{ri}.{n}"))
    w.tail_alpha, w.tail_w, w.tail_b = p("tail_alpha"), p("tail_w"), p("tail_b")
    return w


def make_encoder_weights(folded: Dict[str, np.ndarray]):
    """``snacb_encoder_weights`` struct pointing into ``folded`` (weights.fold_encoder_state_dict)."""
    def p(k):
        a = folded[k]
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], k
        return a.ctypes.data_as(_FP)
    w = EncoderWeights()
    w.conv0_w, w.conv0_b = p("enc.conv0_w"), p("enc.conv0_b")
    for bi in range(4):
        b = w.block[bi]
        b.alpha, b.conv_w, b.conv_b = p(f"enc.b{bi}.alpha"), p(f"enc.b{bi}.conv_w"), p(f"enc.b{bi}.conv_b")
        for ri in range(3):
            for n in ("alpha1", "dw_w", "dw_b", "alpha2", "pw_w", "pw_b"):
                setattr(b.res[ri], n, p(f"enc.b{bi}.r{ri}.{n}"))
    w.final_w, w.final_b = p("enc.final_w"), p("enc.final_b")
    for i in range(3):
        w.in_proj_w[i], w.in_proj_b[i] = p(f"in_proj_w{i}"), p(f"in_proj_b{i}")
        w.codebook[i], w.out_proj_w[i], w.out_proj_b[i] = p(f"codebook{i}"), p(f"out_proj_w{i}"), p(f"out_proj_b{i}")
    return w
