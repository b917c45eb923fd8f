"""Device-side token ingest (include/snacb.h, csrc/kernels_io.cu): the LLM loop's sampled ids -> ready windows.

Replaces, for many streams at once and without a Python int in sight, ``generate_audio_tokens``' SOS/EOS gate
(vllm_inference/modal_audio_stream.py:313-333) and ``stream_audio``'s 28-code buffer policy (:352-396).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

from . import _lib
from .api import SnacbError, SnacDecoder

TOKEN_SOS, TOKEN_EOS = 128257, 128258


class DeviceIngest:
    """``step(tokens[S, n])`` consumes one LLM step of S streams on the GPU and returns the windows that became
    ready: ``(win_tok[W, 28], win_stream[W], tail_tok[Tn, 21], tail_stream[Tn], tail_frames[Tn])`` device tensors
    (views of reused buffers) in (stream, time) order.  The one host read per step is the 8-byte window count."""

    def __init__(self, max_streams: int, device: int = 0):
        import torch
        self._lib = _lib.load()
        self._g = C.c_void_p()
        self.max_streams, self.device = int(max_streams), int(device)
        rc = self._lib.snacb_ingest_create(C.byref(self._g), self.device, self.max_streams)
        if rc != 0:
            self._g = C.c_void_p()
            raise SnacbError(f"snacb_ingest_create failed ({rc})")
        dev = torch.device("cuda", self.device)
        self._tail_tok = torch.zeros((self.max_streams, 21), dtype=torch.int32, device=dev)
        self._tail_stream = torch.zeros(self.max_streams, dtype=torch.int32, device=dev)
        self._tail_frames = torch.zeros(self.max_streams, dtype=torch.int32, device=dev)
        self._counts = torch.zeros(2, dtype=torch.int32, device=dev)
        self._counts_host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self._win_tok = self._win_stream = None

    def close(self):
        if getattr(self, "_g", None) is not None and self._g.value:
            self._lib.snacb_ingest_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, first: int = 0, n: Optional[int] = None):
        import torch
        n = self.max_streams - first if n is None else n
        rc = self._lib.snacb_ingest_reset(self._g, int(first), int(n), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise SnacbError(f"snacb_ingest_reset failed ({rc})")

    def step(self, tokens, n_valid=None, finish=None):
        import torch
        if tokens.dim() == 1:
            tokens = tokens.unsqueeze(1)
        if not tokens.is_cuda or tokens.dtype != torch.int32 or not tokens.is_contiguous():
            raise ValueError("tokens must be a contiguous int32 CUDA tensor [S] or [S, n]")
        S, n = tokens.shape
        if S > self.max_streams:
            raise ValueError("more streams than slots")
        cap = int(self._lib.snacb_ingest_window_capacity(S, n))
        if self._win_tok is None or self._win_tok.shape[0] < cap:
            self._win_tok = torch.zeros((cap, 28), dtype=torch.int32, device=tokens.device)
            self._win_stream = torch.zeros(cap, dtype=torch.int32, device=tokens.device)
        nv = fp = None
        if n_valid is not None:
            if not n_valid.is_cuda or n_valid.dtype != torch.int32 or n_valid.numel() != S:
                raise ValueError("n_valid must be an int32 CUDA tensor [S]")
            nv = n_valid.contiguous().data_ptr()
        if finish is not None:
            if not finish.is_cuda or finish.dtype not in (torch.uint8, torch.bool) or finish.numel() != S:
                raise ValueError("finish must be a uint8 / bool CUDA tensor [S]")
            fp = finish.contiguous().data_ptr()
        st = C.c_void_p(torch.cuda.current_stream(tokens.device).cuda_stream)
        rc = self._lib.snacb_ingest_step(self._g, tokens.data_ptr(), S, n, nv, fp, self._win_tok.data_ptr(),
                                         self._win_stream.data_ptr(), self._win_tok.shape[0], self._tail_tok.data_ptr(),
                                         self._tail_stream.data_ptr(), self._tail_frames.data_ptr(),
                                         self._counts.data_ptr(), st)
        if rc != 0:
            raise SnacbError(f"snacb_ingest_step failed ({rc})")
        self._counts_host.copy_(self._counts, non_blocking=True)
        torch.cuda.current_stream(tokens.device).synchronize()
        nw, nt = int(self._counts_host[0]), int(self._counts_host[1])
        return (self._win_tok[:nw], self._win_stream[:nw], self._tail_tok[:nt], self._tail_stream[:nt],
                self._tail_frames[:nt])

    def state(self) -> Tuple["object", "object"]:
        import numpy as np
        st = np.empty(self.max_streams, dtype=np.int32)
        cnt = np.empty(self.max_streams, dtype=np.int32)
        rc = self._lib.snacb_ingest_state(self._g, st.ctypes.data, cnt.ctypes.data, self.max_streams)
        if rc != 0:
            raise SnacbError(f"snacb_ingest_state failed ({rc})")
        return st, cnt

    def step_decode(self, decoder: SnacDecoder, tokens, n_valid=None, finish=None, seed: Optional[int] = None,
                    precision: str = "fp16") -> List[Tuple[int, "object"]]:
        """One LLM step end to end on the device: ingest, then ONE batched decode of every window that became ready
        (plus one small decode per remainder length at end of stream).  Returns [(stream, int16 PCM device tensor)]
        in (stream, time) order -- what ``stream_audio`` yields, for all streams.
        ``seed=None`` (default): every call draws fresh NoiseBlock noise, as the reference does with ``torch.randn`` on
        every decode (a constant seed would repeat one noise sequence in every window of every step)."""
        import torch
        if seed is None:
            self._seed = getattr(self, "_seed", 0) + 1
            seed = self._seed
        wt, ws, tt, ts, tf = self.step(tokens, n_valid, finish)
        out = []                                   # (stream, order within the stream, pcm)
        if wt.shape[0]:
            pcm = decoder.decode(wt, raw_ids=True, seed=seed, precision=precision)
            out += [(s, i, pcm[i]) for i, s in enumerate(ws.cpu().tolist())]      # already in (stream, time) order
        if tt.shape[0]:
            frames, ids = tf.cpu().tolist(), ts.cpu().tolist()
            for fr in sorted(set(frames)):
                rows = [i for i, f in enumerate(frames) if f == fr]
                sel = tt[torch.tensor(rows, device=tt.device)][:, : 7 * fr].contiguous()
                pcm = decoder.decode(sel, raw_ids=True, seed=seed, precision=precision)
                out += [(ids[r], 1 << 30, pcm[j]) for j, r in enumerate(rows)]    # a remainder follows its stream's windows
        out.sort(key=lambda e: (e[0], e[1]))
        return [(s, p) for (s, _, p) in out]
