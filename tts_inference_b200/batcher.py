"""Python front of the C++ window batcher (include/snacb.h, csrc/batcher.cpp).

Replaces ``stream_audio``'s one-stream-at-a-time buffer policy
(vllm_inference/modal_audio_stream.py:352-396): producers ``push`` token ids of many streams,
``flush`` decodes every ready window of every stream in one batched launch sequence, straight
into a pinned host buffer.  ``push`` is thread-safe (streams are sharded over independently
locked tables); the flush calls belong to one flusher thread.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from .api import SnacbError, SnacDecoder

POLICY_CHUNK = 0     # shipped stream_audio rule: non-overlapping 28-code chunks, all samples emitted
POLICY_SLIDING = 1   # constants' rule (modal_audio_stream.py:86-95): last 28 every 7, emit [2048:4096]


class WindowBatcher:
    def __init__(self, decoder: Optional[SnacDecoder], policy: int = POLICY_CHUNK, raw_ids: bool = True,
                 max_windows: int = 1024, precision: str = "fp16"):
        """``decoder=None`` builds a queue-only batcher (``push`` / ``end`` / ``take``; no GPU needed)."""
        self._lib = _lib.load()
        self._dec = decoder
        self._b = C.c_void_p()
        self.max_windows = int(max_windows)
        self.policy = int(policy)
        flags = (_lib.RAW_IDS if raw_ids else 0) | (_lib.FP32 if precision == "fp32" else 0) | \
            (_lib.BF16 if precision == "bf16" else 0)
        if decoder is not None:
            decoder._adopt(self)                            # closed with (before) its decoder
        rc = self._lib.snacb_batcher_create(C.byref(self._b), decoder._h if decoder is not None else None,
                                            self.policy, flags, self.max_windows)
        if rc != 0:
            raise SnacbError(f"snacb_batcher_create failed ({rc})")
        self._seed = 0
        self._slots = []          # two output slots for the pipelined flush: (ids, off, len, pcm numpy view, keepalive)
        self._inflight = []       # (slot index, n chunks) of submitted, not yet waited flushes

    def _slot(self, i: int):
        while len(self._slots) <= i:
            per = 2048 if self.policy == POLICY_SLIDING else 8192
            n = self.max_windows * per
            keep = None
            try:                  # pinned memory: the device->host copy of a flush is then an asynchronous DMA
                import torch
                keep = torch.empty(n, dtype=torch.int16).pin_memory() if torch.cuda.is_available() else None
            except Exception:     # noqa: BLE001
                keep = None
            pcm = keep.numpy() if keep is not None else np.empty(n, dtype=np.int16)
            self._slots.append((np.empty(self.max_windows, dtype=np.uint64), np.empty(self.max_windows, dtype=np.int64),
                                np.empty(self.max_windows, dtype=np.int32), pcm, keep))
        return self._slots[i]

    def close(self):
        if self._b.value:
            self._lib.snacb_batcher_destroy(self._b)
            self._b = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push(self, stream_id: int, tokens) -> None:
        t = np.ascontiguousarray(tokens, dtype=np.int32).reshape(-1)
        rc = self._lib.snacb_batcher_push(self._b, C.c_uint64(stream_id), t.ctypes.data, t.size)
        if rc != 0:
            raise SnacbError(f"snacb_batcher_push failed ({rc})" + (": the stream has ended" if rc == -4 else ""))

    def end(self, stream_id: int) -> None:
        rc = self._lib.snacb_batcher_end(self._b, C.c_uint64(stream_id))
        if rc != 0:
            raise SnacbError(f"snacb_batcher_end failed ({rc})")

    def forget(self, stream_id: int) -> None:
        """Release an ended stream's id (until then a push to it is refused rather than starting a new stream)."""
        rc = self._lib.snacb_batcher_forget(self._b, C.c_uint64(stream_id))
        if rc != 0:
            raise SnacbError(f"snacb_batcher_forget failed ({rc})")

    def pending(self) -> int:
        return int(self._lib.snacb_batcher_pending(self._b))

    def take(self, max_windows: Optional[int] = None) -> List[Tuple[int, np.ndarray]]:
        """Pop ready windows without decoding: [(stream_id, int32 tokens of 7 * frames entries)]."""
        m = self.max_windows if max_windows is None else int(max_windows)
        ids = np.empty(m, dtype=np.uint64); fr = np.empty(m, dtype=np.int32); tok = np.empty((m, 28), dtype=np.int32)
        n = self._lib.snacb_batcher_take(self._b, m, ids.ctypes.data, fr.ctypes.data, tok.ctypes.data)
        if n < 0:
            raise SnacbError(f"snacb_batcher_take failed ({n})")
        return [(int(ids[i]), tok[i, : 7 * fr[i]].copy()) for i in range(n)]

    def _next_seed(self, seed: Optional[int]) -> int:
        # fresh NoiseBlock noise on every flush, as the reference draws torch.randn per decode
        if seed is None:
            self._seed += 1
            return self._seed
        return int(seed)

    def _collect(self, slot: int, n: int) -> List[Tuple[int, np.ndarray]]:
        ids, off, ln, pcm, _ = self._slot(slot)
        return [(int(ids[i]), pcm[off[i]: off[i] + ln[i]].copy()) for i in range(n)]

    def _fail(self, what: str, rc: int):
        msg = self._lib.snacb_last_error(self._dec._h) if self._dec is not None else b""
        raise SnacbError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def flush(self, seed: Optional[int] = None) -> List[Tuple[int, np.ndarray]]:
        """Decode all ready windows; returns [(stream_id, int16 samples)], each stream's chunks in time order."""
        if self._inflight:
            raise SnacbError("flush() while pipelined flushes are outstanding: call flush_wait() first")
        ids, off, ln, pcm, _ = self._slot(0)
        n = self._lib.snacb_batcher_flush(self._b, C.c_uint64(self._next_seed(seed)), self.max_windows, ids.ctypes.data,
                                          off.ctypes.data, ln.ctypes.data, pcm.ctypes.data, pcm.size)
        if n < 0:
            self._fail("snacb_batcher_flush", n)
        return self._collect(0, n)

    def flush_submit(self, seed: Optional[int] = None) -> int:
        """Queue the decode of everything that is ready and return (at most two outstanding); ``flush_wait`` returns
        the oldest one's chunks.  submit(tick i + 1) followed by wait() (tick i) overlaps copy-out with decode."""
        if len(self._inflight) >= 2:
            raise SnacbError("two flushes outstanding: call flush_wait() first")
        slot = 0 if not self._inflight else 1 - self._inflight[-1][0]
        ids, off, ln, pcm, _ = self._slot(slot)
        n = self._lib.snacb_batcher_flush_submit(self._b, C.c_uint64(self._next_seed(seed)), self.max_windows,
                                                 ids.ctypes.data, off.ctypes.data, ln.ctypes.data, pcm.ctypes.data, pcm.size)
        if n < 0:
            self._fail("snacb_batcher_flush_submit", n)
        if n > 0:
            self._inflight.append((slot, n))
        return n

    def flush_wait(self) -> List[Tuple[int, np.ndarray]]:
        if not self._inflight:
            return []
        slot, n = self._inflight.pop(0)
        rc = self._lib.snacb_batcher_flush_wait(self._b)
        if rc != 0:
            self._fail("snacb_batcher_flush_wait", rc)
        return self._collect(slot, n)
