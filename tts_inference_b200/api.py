"""Batched Python front of libsnacb: device tensors in, device tensors out, no hidden sync.

PyTorch is used only for device memory and streams; every FLOP of the path runs in the
hand-written sm_100a kernels behind the C ABI (include/snacb.h).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .weights import fold_state_dict

FRAME = 7
WINDOW = 28


class SnacbError(RuntimeError):
    pass


class SnacDecoder:
    """One handle = one GPU.  Replaces the module-global ``snac_model`` the reference's helper
    closes over (vllm_inference/modal_audio_stream.py:79-80,106-129)."""

    def __init__(self, state_dict: Mapping[str, object], device: int = 0, folded: bool = False):
        """``state_dict``: the checkpoint's tensors under their upstream names (weight-norm folded here), or -- with
        ``folded=True`` -- the output of ``weights.fold_state_dict`` / ``weights.load_folded``."""
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.device = int(device)
        folded = dict(state_dict) if folded else fold_state_dict(state_dict)
        w = _lib.make_weights(folded)
        rc = self._lib.snacb_create(C.byref(self._h), C.byref(w), self.device)
        if rc != 0:
            msg = self._lib.snacb_last_error(None)
            self._h = C.c_void_p()
            raise SnacbError(f"snacb_create failed ({rc}): {msg.decode() if msg else ''}")

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            for child in list(getattr(self, "_children", ())):      # sessions / streamers hold pointers into the handle
                child.close()
            self._lib.snacb_destroy(self._h)
            self._h = C.c_void_p()

    def _adopt(self, child):
        import weakref
        if not hasattr(self, "_children"):
            self._children = weakref.WeakSet()
        self._children.add(child)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self._lib.snacb_last_error(self._h)
            raise SnacbError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def _stream_ptr(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)      # the handle's GPU, not the current one

    def chain_modes(self):
        """Per DecoderBlock: 0 per-layer kernels, 1 general fused chain, 2 alpha-folded fused chain (fp16 operands);
        decided from the checkpoint's Snake alphas at load (include/snacb.h snacb_chain_modes)."""
        m = (C.c_int32 * 4)()
        self._check(self._lib.snacb_chain_modes(self._h, m), "snacb_chain_modes")
        return list(m)

    @staticmethod
    def _flags(raw_ids, extract_slice, precision, keep_taps=False, stream_fp32=False, unfused=False) -> int:
        if precision not in ("fp16", "bf16", "fp32"):
            raise ValueError("precision must be 'fp16', 'bf16' or 'fp32'")
        f = 0
        f |= _lib.RAW_IDS if raw_ids else 0
        f |= _lib.EXTRACT_SLICE if extract_slice else 0
        f |= _lib.FP32 if precision == "fp32" else 0
        f |= _lib.BF16 if precision == "bf16" else 0
        f |= _lib.KEEP_TAPS if keep_taps else 0
        f |= _lib.STREAM_FP32 if stream_fp32 else 0
        f |= _lib.UNFUSED if unfused else 0
        return f

    def samples_out(self, frames: int, extract_slice: bool) -> int:
        return int(self._lib.snacb_samples_out(int(frames), _lib.EXTRACT_SLICE if extract_slice else 0))

    def set_group_bytes(self, nbytes: int):
        self._check(self._lib.snacb_set_group_bytes(self._h, int(nbytes)), "snacb_set_group_bytes")

    def stats(self) -> Tuple[int, int]:
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._lib.snacb_stats(self._h, C.byref(a), C.byref(b)), "snacb_stats")
        return int(a.value), int(b.value)

    # ------------------------------------------------------------------ device API
    def unpack(self, tokens, raw_ids: bool = False):
        """tokens: cuda int32 [B, n] -> (c0 [B,F], c1 [B,2F], c2 [B,4F]) int32, F = n // 7."""
        import torch
        assert tokens.is_cuda and tokens.dtype == torch.int32 and tokens.dim() == 2 and tokens.is_contiguous()
        B, n = tokens.shape
        F_ = n // FRAME
        c0 = torch.empty((B, F_), dtype=torch.int32, device=tokens.device)
        c1 = torch.empty((B, 2 * F_), dtype=torch.int32, device=tokens.device)
        c2 = torch.empty((B, 4 * F_), dtype=torch.int32, device=tokens.device)
        rc = self._lib.snacb_unpack(self._h, tokens.data_ptr(), B, n, _lib.RAW_IDS if raw_ids else 0,
                                    c0.data_ptr(), c1.data_ptr(), c2.data_ptr(), self._stream_ptr())
        self._check(rc, "snacb_unpack")
        return c0, c1, c2

    def decode(self, tokens, *, raw_ids: bool = False, extract_slice: bool = False,
               noise: Optional[Sequence] = None, seed: int = 0, precision: str = "fp16",
               out=None, return_wave: bool = False, keep_taps: bool = False, stream_fp32: bool = False,
               unfused: bool = False, stream_keys=None, sample_range: Optional[Tuple[int, int]] = None):
        """tokens: cuda int32 [B, n>=7F] (trailing partial frame ignored, as the helper does).
        Returns int16 [B, samples] (and the fp32 waveform when ``return_wave``).
        ``stream_keys`` (cuda int32 [B], optional): key of each row's built-in noise instead of its position in the batch.
        ``sample_range`` (lo, hi): write only samples [lo, hi) of every row and compute only their receptive field."""
        import torch
        assert tokens.is_cuda and tokens.dtype == torch.int32 and tokens.dim() == 2 and tokens.is_contiguous()
        B, n = tokens.shape
        frames = n // FRAME
        flags = self._flags(raw_ids, extract_slice, precision, keep_taps, stream_fp32, unfused)
        ns = self.samples_out(frames, extract_slice)
        lo = hi = 0
        if sample_range is not None:
            lo, hi = int(sample_range[0]), int(sample_range[1])
            if not 0 <= lo < hi <= 2048 * frames:
                raise ValueError("sample_range outside the decoded samples")
            ns = hi - lo
        if out is None:
            out = torch.empty((B, ns), dtype=torch.int16, device=tokens.device)
        else:
            assert out.is_cuda and out.dtype == torch.int16 and out.is_contiguous() and out.numel() == B * ns
        wave = torch.empty((B, ns), dtype=torch.float32, device=tokens.device) if return_wave else None
        nz_arr = None
        keep = []
        if noise is not None:
            t0 = 4 * frames
            lens = [t0 * 8, t0 * 64, t0 * 256, t0 * 512]
            nz_arr = (C.c_void_p * 4)()
            for i, t in enumerate(noise):
                t = t.reshape(B, -1)
                assert t.is_cuda and t.dtype == torch.float32 and t.shape[1] == lens[i], (t.shape, lens[i])
                t = t.contiguous()
                keep.append(t)
                nz_arr[i] = t.data_ptr()
        keys = None
        if stream_keys is not None:
            assert stream_keys.is_cuda and stream_keys.dtype == torch.int32 and stream_keys.numel() == B
            stream_keys = stream_keys.contiguous()
            keys = stream_keys.data_ptr()
        rc = self._lib.snacb_decode_range(self._h, tokens.data_ptr(), B, n, frames, flags, nz_arr, C.c_uint64(seed), keys,
                                          lo, hi, out.data_ptr(), wave.data_ptr() if wave is not None else None,
                                          self._stream_ptr())
        self._check(rc, "snacb_decode")
        return (out, wave) if return_wave else out

    def decode_windows(self, tokens, **kw):
        """[B, 28] sliding windows -> int16 [B, 2048] (samples [2048:4096] of each decode)."""
        assert tokens.shape[1] == WINDOW
        kw.setdefault("extract_slice", True)
        return self.decode(tokens, **kw)

    def decode_full(self, tokens, **kw):
        """[B, 7F] whole utterances -> int16 [B, 2048 F]."""
        kw["extract_slice"] = False
        return self.decode(tokens, **kw)

    # ------------------------------------------------------------------ host API (what the helper's caller sees)
    def decode_host(self, tokens: np.ndarray, *, raw_ids: bool = False, extract_slice: bool = False,
                    seed: int = 0, precision: str = "fp16", out: Optional[np.ndarray] = None) -> np.ndarray:
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        assert tokens.ndim == 2
        B, n = tokens.shape
        frames = n // FRAME
        ns = self.samples_out(frames, extract_slice)
        if out is None:
            out = np.empty((B, ns), dtype=np.int16)
        assert out.dtype == np.int16 and out.flags["C_CONTIGUOUS"] and out.size == B * ns
        rc = self._lib.snacb_decode_host(self._h, tokens.ctypes.data, B, n, frames,
                                         self._flags(raw_ids, extract_slice, precision), C.c_uint64(seed),
                                         out.ctypes.data)
        self._check(rc, "snacb_decode_host")
        return out

    def decode_host_ptr(self, tok_ptr: int, B: int, n: int, pcm_ptr: int, *, raw_ids=False, extract_slice=False,
                        seed: int = 0, precision: str = "fp16"):
        """Same, raw host pointers (e.g. pinned torch tensors) -- used by the benchmark's e2e leg."""
        rc = self._lib.snacb_decode_host(self._h, tok_ptr, B, n, n // FRAME,
                                         self._flags(raw_ids, extract_slice, precision), C.c_uint64(seed), pcm_ptr)
        self._check(rc, "snacb_decode_host")

    def submit_host_ptr(self, tok_ptr: int, B: int, n: int, pcm_ptr: int, *, raw_ids=False, extract_slice=False,
                        seed: int = 0, precision: str = "fp16"):
        """Pipelined host boundary: queue copy-in + decode + copy-out of one step and return (at most two outstanding).
        ``wait_host()`` blocks until the oldest outstanding step's PCM is in its host buffer."""
        rc = self._lib.snacb_decode_host_submit(self._h, tok_ptr, B, n, n // FRAME,
                                                self._flags(raw_ids, extract_slice, precision), C.c_uint64(seed), pcm_ptr)
        self._check(rc, "snacb_decode_host_submit")

    def wait_host(self):
        self._check(self._lib.snacb_decode_host_wait(self._h), "snacb_decode_host_wait")

    # ------------------------------------------------------------------ stateful streaming
    def open_session(self, n_slots: int, window_frames: int = 32, *, raw_ids: bool = True, precision: str = "fp16") -> "StreamingSession":
        """Incremental decode of growing streams with per-slot, per-stage state in HBM (include/snacb.h, snacb_session_*).
        ``window_frames``: frames of activations every slot keeps (a sliding window; streams are unbounded)."""
        return StreamingSession(self, n_slots, window_frames, raw_ids=raw_ids, precision=precision)

    # ------------------------------------------------------------------ per-stage timing
    def profile(self, enable: bool = True):
        self._check(self._lib.snacb_profile(self._h, 1 if enable else 0), "snacb_profile")

    def profile_report(self) -> Dict[str, Tuple[int, float]]:
        """{stage: (launches, total_ms)} since profile(True); synchronises the device."""
        buf = C.create_string_buffer(1 << 16)
        self._check(self._lib.snacb_profile_report(self._h, buf, len(buf)), "snacb_profile_report")
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split()
            out[name] = (int(cnt), float(ms))
        return out

    # ------------------------------------------------------------------ debug taps
    def taps(self) -> Dict[str, np.ndarray]:
        out: Dict[str, np.ndarray] = {}
        n = self._lib.snacb_debug_tap_count(self._h)
        for i in range(max(n, 0)):
            name = C.create_string_buffer(64)
            rows, cols = C.c_int64(), C.c_int64()
            self._check(self._lib.snacb_debug_tap_info(self._h, i, name, 64, C.byref(rows), C.byref(cols)), "tap_info")
            a = np.empty((rows.value, cols.value), dtype=np.float32)
            self._check(self._lib.snacb_debug_tap_copy(self._h, i, a.ctypes.data, a.size), "tap_copy")
            out[name.value.decode()] = a
        return out


class StreamingSession:
    """``snacb_session``: slots 0..n_slots-1 each hold one growing stream; ``step`` appends frames to a contiguous range of
    slots that are at the same position and returns the samples that became final -- no recompute of the prefix, and the
    concatenation of all steps of a slot equals ``SnacDecoder.decode(..., stream_keys=key)`` of the finished stream bit for
    bit (tests/test_gpu_api.py).  Each slot keeps a sliding window of ``max_frames`` frames of activations; a stream may be
    arbitrarily long, a single step may add at most ``max_frames - 16`` frames to a non-empty window."""

    def __init__(self, decoder: SnacDecoder, n_slots: int, max_frames: int, *, raw_ids: bool = True, precision: str = "fp16"):
        if precision not in ("fp16", "bf16"):
            raise ValueError("a session runs the tensor-core path: precision 'fp16' or 'bf16'")
        self._dec, self._lib = decoder, decoder._lib
        self._s = C.c_void_p()
        flags = (_lib.RAW_IDS if raw_ids else 0) | (_lib.BF16 if precision == "bf16" else 0)
        rc = self._lib.snacb_session_create(decoder._h, int(n_slots), int(max_frames), flags, C.byref(self._s))
        decoder._check(rc, "snacb_session_create")
        self.n_slots = int(n_slots)
        self.max_frames = int(self._lib.snacb_session_max_frames(self._s))
        decoder._adopt(self)                                # closed with (before) its decoder

    def close(self):
        if getattr(self, "_s", None) is not None and self._s.value:
            self._lib.snacb_session_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nbytes(self) -> int:
        return int(self._lib.snacb_session_bytes(self._s))

    def frames(self, slot: int) -> int:
        return int(self._lib.snacb_session_frames(self._s, int(slot)))

    def emitted(self, slot: int) -> int:
        return int(self._lib.snacb_session_emitted(self._s, int(slot)))

    def reset(self, slot0: int = 0, n: Optional[int] = None):
        n = self.n_slots - slot0 if n is None else n
        self._dec._check(self._lib.snacb_session_reset(self._s, int(slot0), int(n)), "snacb_session_reset")

    def next_emit(self, slot: int, new_frames: int, final: bool = False) -> int:
        r = int(self._lib.snacb_session_next_emit(self._s, int(slot), int(new_frames), 1 if final else 0))
        if r < 0:
            raise SnacbError(f"snacb_session_next_emit failed ({r})")
        return r

    def step_multi(self, slots, new_tokens, *, seed: int = 0, stream_keys=None):
        """One step for an arbitrary set of slots that may be at different positions of their streams (asynchronous
        streams): ``slots`` a sequence of n distinct slot indices, ``new_tokens`` cuda int32 [n, 7k] with k >= 1 new frames
        for each.  Returns int16 [n, m], the same m for all (2048 k once a stream holds 3 frames).  Not for end of stream."""
        import torch
        slots = np.ascontiguousarray(np.asarray(slots, dtype=np.int32))
        assert new_tokens.is_cuda and new_tokens.dtype == torch.int32 and new_tokens.dim() == 2
        new_tokens = new_tokens.contiguous()
        n, w = new_tokens.shape
        assert slots.shape == (n,) and w >= FRAME
        k = w // FRAME
        m = max(self.next_emit(int(s), k) for s in slots[:1]) if n else 0
        out = torch.empty((n, m), dtype=torch.int16, device=new_tokens.device)
        keys = None
        if stream_keys is not None:
            assert stream_keys.is_cuda and stream_keys.dtype == torch.int32 and stream_keys.numel() == n
            stream_keys = stream_keys.contiguous()
            keys = stream_keys.data_ptr()
        got = C.c_int(0)
        rc = self._lib.snacb_session_step_multi(self._s, n, slots.ctypes.data, new_tokens.data_ptr(), w, k, C.c_uint64(seed), keys,
                                                out.data_ptr() if m else None, m, C.byref(got), self._dec._stream_ptr())
        self._dec._check(rc, "snacb_session_step_multi")
        assert got.value == m, (got.value, m)
        return out

    def step(self, slot0: int, new_tokens, *, final: bool = False, seed: int = 0, stream_keys=None, out=None):
        """new_tokens: cuda int32 [n, 7k] (k >= 0 new frames for slots slot0 .. slot0+n-1; k = 0 only with ``final``).
        Returns int16 [n, m]: the m samples per slot that became final (m may be 0)."""
        import torch
        assert new_tokens.is_cuda and new_tokens.dtype == torch.int32 and new_tokens.dim() == 2
        new_tokens = new_tokens.contiguous()
        n, w = new_tokens.shape
        k = w // FRAME
        m = self.next_emit(slot0, k, final)
        if out is None:
            out = torch.empty((n, m), dtype=torch.int16, device=new_tokens.device)
        else:
            assert out.is_cuda and out.dtype == torch.int16 and out.is_contiguous() and out.shape == (n, m)
        keys = None
        if stream_keys is not None:
            assert stream_keys.is_cuda and stream_keys.dtype == torch.int32 and stream_keys.numel() == n
            stream_keys = stream_keys.contiguous()
            keys = stream_keys.data_ptr()
        got = C.c_int(0)
        rc = self._lib.snacb_session_step(self._s, int(slot0), n, new_tokens.data_ptr() if w else None, w, k, 1 if final else 0,
                                          C.c_uint64(seed), keys, out.data_ptr() if m else None, m, C.byref(got),
                                          self._dec._stream_ptr())
        self._dec._check(rc, "snacb_session_step")
        assert got.value == m, (got.value, m)
        return out
