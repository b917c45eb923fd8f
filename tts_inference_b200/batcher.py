"""Python front of the C++ window batcher (include/snacb.h, csrc/batcher.cpp).

Replaces ``stream_audio``'s one-stream-at-a-time buffer policy
(vllm_inference/modal_audio_stream.py:352-396): producers ``push`` token ids of many streams,
``flush`` decodes every ready window of every stream in one batched launch sequence.
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
    def __init__(self, decoder: SnacDecoder, policy: int = POLICY_CHUNK, raw_ids: bool = True,
                 max_windows: int = 1024, precision: str = "fp16"):
        self._lib = _lib.load()
        self._dec = decoder
        self._b = C.c_void_p()
        self.max_windows = int(max_windows)
        flags = (_lib.RAW_IDS if raw_ids else 0) | (_lib.FP32 if precision == "fp32" else 0) | \
            (_lib.BF16 if precision == "bf16" else 0)
        rc = self._lib.snacb_batcher_create(C.byref(self._b), decoder._h, int(policy), flags, self.max_windows)
        if rc != 0:
            raise SnacbError(f"snacb_batcher_create failed ({rc})")
        self._ids = np.empty(self.max_windows, dtype=np.uint64)
        self._off = np.empty(self.max_windows, dtype=np.int64)
        self._len = np.empty(self.max_windows, dtype=np.int32)
        self._pcm = np.empty(self.max_windows * 8192, dtype=np.int16)

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
            raise SnacbError(f"snacb_batcher_push failed ({rc})")

    def end(self, stream_id: int) -> None:
        rc = self._lib.snacb_batcher_end(self._b, C.c_uint64(stream_id))
        if rc != 0:
            raise SnacbError(f"snacb_batcher_end failed ({rc})")

    def pending(self) -> int:
        return int(self._lib.snacb_batcher_pending(self._b))

    def flush(self, seed: Optional[int] = None) -> List[Tuple[int, np.ndarray]]:
        """Decode all ready windows; returns [(stream_id, int16 samples)] in queue order.
        ``seed=None`` (default): fresh NoiseBlock noise on every flush, as the reference draws ``torch.randn`` per decode."""
        if seed is None:
            self._seed = getattr(self, "_seed", 0) + 1
            seed = self._seed
        n = self._lib.snacb_batcher_flush(self._b, C.c_uint64(seed), self.max_windows, self._ids.ctypes.data,
                                          self._off.ctypes.data, self._len.ctypes.data, self._pcm.ctypes.data,
                                          self._pcm.size)
        if n < 0:
            msg = self._lib.snacb_last_error(self._dec._h)
            raise SnacbError(f"snacb_batcher_flush failed ({n}): {msg.decode() if msg else ''}")
        return [(int(self._ids[i]), self._pcm[self._off[i]: self._off[i] + self._len[i]].copy()) for i in range(n)]
