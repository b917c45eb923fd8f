"""Egress formats on the device (include/snacb.h, csrc/kernels_io.cu): int16 PCM -> the bytes the reference's
endpoints send -- base64 text per chunk (``/ws/audio``, vllm_inference/modal_audio_stream.py:483-487) and RIFF/WAVE
files (``/generate``, ``/generate-batch``, :561-566, :650-657).  Byte-exact against ``base64.b64encode`` / ``wave``."""
from __future__ import annotations

import ctypes as C

from . import _lib
from .api import SnacbError

SAMPLE_RATE = 24000


def _prep(pcm):
    import torch
    if not pcm.is_cuda or pcm.dtype != torch.int16:
        raise ValueError("pcm must be an int16 CUDA tensor [n, samples]")
    if pcm.dim() == 1:
        pcm = pcm.unsqueeze(0)
    return pcm.contiguous(), C.c_void_p(torch.cuda.current_stream(pcm.device).cuda_stream)


def pcm_to_base64(pcm, out=None):
    """[n, samples] int16 -> [n, 4*ceil(2*samples/3)] uint8 ASCII, one independent base64 string per row."""
    import torch
    lib = _lib.load()
    pcm, st = _prep(pcm)
    n, samples = pcm.shape
    ln = int(lib.snacb_base64_len(2 * samples))
    if out is None:
        out = torch.empty((n, ln), dtype=torch.uint8, device=pcm.device)
    elif not out.is_cuda or out.dtype != torch.uint8 or out.numel() != n * ln or not out.is_contiguous():
        raise ValueError("out must be a contiguous uint8 CUDA tensor of n * base64_len bytes")
    rc = lib.snacb_pcm_to_base64(pcm.data_ptr(), n, samples, out.data_ptr(), st)
    if rc != 0:
        raise SnacbError(f"snacb_pcm_to_base64 failed ({rc})")
    return out


def pcm_to_wav(pcm, sample_rate: int = SAMPLE_RATE, out=None):
    """[n, samples] int16 -> [n, 44 + 2*samples] uint8: one mono 16-bit RIFF/WAVE file per row."""
    import torch
    lib = _lib.load()
    pcm, st = _prep(pcm)
    n, samples = pcm.shape
    ln = 44 + 2 * samples
    if out is None:
        out = torch.empty((n, ln), dtype=torch.uint8, device=pcm.device)
    elif not out.is_cuda or out.dtype != torch.uint8 or out.numel() != n * ln or not out.is_contiguous():
        raise ValueError("out must be a contiguous uint8 CUDA tensor of n * (44 + 2*samples) bytes")
    rc = lib.snacb_pcm_to_wav(pcm.data_ptr(), n, samples, int(sample_rate), out.data_ptr(), st)
    if rc != 0:
        raise SnacbError(f"snacb_pcm_to_wav failed ({rc})")
    return out
