"""``StreamServer``: the native multi-stream front of the stateful session (``snacb_streamer_*``, include/snacb.h) -- token
ids in from any number of producer threads, each stream's newly final int16 samples out per tick.  The stateful counterpart
of ``WindowBatcher``; replaces the per-stream Python buffering of ``stream_audio`` (modal_audio_stream.py:340-409)."""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

from . import _lib
from .api import SnacDecoder, SnacbError


class StreamServer:
    def __init__(self, decoder: SnacDecoder, max_streams: int, window_frames: int = 32, *, raw_ids: bool = True,
                 precision: str = "fp16", min_frames: int = 1, max_samples_per_tick: int = 0):
        if precision not in ("fp16", "bf16"):
            raise ValueError("precision 'fp16' or 'bf16'")
        self._dec, self._lib = decoder, decoder._lib
        self._s = C.c_void_p()
        flags = (_lib.RAW_IDS if raw_ids else 0) | (_lib.BF16 if precision == "bf16" else 0)
        rc = self._lib.snacb_streamer_create(C.byref(self._s), decoder._h, int(max_streams), int(window_frames), flags, int(min_frames))
        decoder._check(rc, "snacb_streamer_create")
        self.max_streams = int(max_streams)
        decoder._adopt(self)                                # closed with (before) its decoder
        import torch
        cap = max_samples_per_tick or self.max_streams * 2048 * 24
        self._pcm = torch.empty(cap, dtype=torch.int16).pin_memory()        # pinned: the copy-out is a DMA
        self._ids = np.empty(self.max_streams, dtype=np.uint64)
        self._off = np.empty(self.max_streams, dtype=np.int64)
        self._len = np.empty(self.max_streams, dtype=np.int32)
        self._seed = 0

    def close(self):
        if getattr(self, "_s", None) is not None and self._s.value:
            self._lib.snacb_streamer_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push(self, stream_id: int, ids) -> None:
        a = np.ascontiguousarray(np.clip(np.asarray(ids, dtype=np.int64), -(2 ** 31), 2 ** 31 - 1).astype(np.int32))
        rc = self._lib.snacb_streamer_push(self._s, C.c_uint64(stream_id), a.ctypes.data, a.size)
        if rc != 0:
            raise SnacbError(f"snacb_streamer_push failed ({rc}): no free slot, or the stream has ended")

    def end(self, stream_id: int) -> None:
        rc = self._lib.snacb_streamer_end(self._s, C.c_uint64(stream_id))
        if rc != 0:
            raise SnacbError(f"snacb_streamer_end failed ({rc}): unknown stream")

    def active(self) -> int:
        return int(self._lib.snacb_streamer_active(self._s))

    def tick(self, seed: int = 0) -> List[Tuple[int, np.ndarray]]:
        """Serve every stream that is due; returns [(stream id, new int16 samples)] (copies)."""
        n = self._lib.snacb_streamer_tick(self._s, C.c_uint64(seed), self.max_streams, self._ids.ctypes.data, self._off.ctypes.data,
                                          self._len.ctypes.data, self._pcm.data_ptr(), self._pcm.numel())
        if n < 0:
            msg = self._lib.snacb_last_error(self._dec._h)
            raise SnacbError(f"snacb_streamer_tick failed ({n}): {msg.decode() if msg else ''}")
        pcm = self._pcm.numpy()
        return [(int(self._ids[i]), pcm[self._off[i]: self._off[i] + self._len[i]].copy()) for i in range(n)]
