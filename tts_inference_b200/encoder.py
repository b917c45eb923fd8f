"""SNAC encode path: audio -> codes -> token ids, behind ``snacb_encode`` / ``snacb_pack_tokens``
(include/snacb.h).  Upstream ``snac.SNAC.encode``; the reference only decodes at inference, this is the other half of the
codec for dataset tokenisation and round-trip checks.  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Tuple

from . import _lib
from .api import SnacbError
from .weights import fold_encoder_state_dict


class SnacEncoder:
    def __init__(self, state_dict: Mapping[str, object], device: int = 0, folded: bool = False):
        self._lib = _lib.load()
        self._e = C.c_void_p()
        self.device = int(device)
        folded = dict(state_dict) if folded else fold_encoder_state_dict(state_dict)
        w = _lib.make_encoder_weights(folded)
        rc = self._lib.snacb_encoder_create(C.byref(self._e), C.byref(w), self.device)
        if rc != 0:
            msg = self._lib.snacb_encoder_last_error(None)
            self._e = C.c_void_p()
            raise SnacbError(f"snacb_encoder_create failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_e", None) is not None and self._e.value:
            self._lib.snacb_encoder_destroy(self._e)
            self._e = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launches(self) -> int:
        return int(self._lib.snacb_encoder_launches(self._e))

    def encode(self, audio, *, return_latent: bool = False, return_dist: bool = False):
        """audio: cuda float32 [B, n] -> (c0 [B, F], c1 [B, 2F], c2 [B, 4F]) int32, F = ceil(n / 2048) (the tail of the last
        frame is zero padding, as ``SNAC.preprocess`` pads).  Optionally also the latent z [B, 4F, 768] and the winning
        distance of every code [B, 7F] (level by level)."""
        import torch
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 2
        audio = audio.contiguous()
        B, n = audio.shape
        F_ = int(self._lib.snacb_encode_frames(n))
        dev = audio.device
        c = [torch.empty((B, F_ << l), dtype=torch.int32, device=dev) for l in range(3)]
        z = torch.empty((B, 4 * F_, 768), dtype=torch.float32, device=dev) if return_latent else None
        d = torch.empty((7 * F_ * B,), dtype=torch.float32, device=dev) if return_dist else None
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = self._lib.snacb_encode(self._e, audio.data_ptr(), B, n, n, c[0].data_ptr(), c[1].data_ptr(), c[2].data_ptr(),
                                    z.data_ptr() if z is not None else None, d.data_ptr() if d is not None else None, st)
        if rc != 0:
            msg = self._lib.snacb_encoder_last_error(self._e)
            raise SnacbError(f"snacb_encode failed ({rc}): {msg.decode() if msg else ''}")
        out = (c[0], c[1], c[2])
        if return_latent:
            out = out + (z,)
        if return_dist:
            ds = [d[: B * F_].view(B, F_), d[B * F_: 3 * B * F_].view(B, 2 * F_), d[3 * B * F_:].view(B, 4 * F_)]
            out = out + (ds,)
        return out

    def pack_tokens(self, c0, c1, c2, raw_ids: bool = True):
        """codes -> token ids [B, 7F] int32 (inverse of ``SnacDecoder.unpack``)."""
        import torch
        B, F_ = c0.shape
        tok = torch.empty((B, 7 * F_), dtype=torch.int32, device=c0.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = self._lib.snacb_pack_tokens(c0.contiguous().data_ptr(), c1.contiguous().data_ptr(), c2.contiguous().data_ptr(), B, F_,
                                         _lib.RAW_IDS if raw_ids else 0, tok.data_ptr(), st)
        if rc != 0:
            raise SnacbError(f"snacb_pack_tokens failed ({rc})")
        return tok

    def encode_tokens(self, audio, raw_ids: bool = True):
        c0, c1, c2 = self.encode(audio)
        return self.pack_tokens(c0, c1, c2, raw_ids=raw_ids)
