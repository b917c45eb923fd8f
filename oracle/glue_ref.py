"""Restatement of the reference's integer glue around ``snac.decode`` (ORACLE).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  This part of the path IS pinned:
``tests/golden/make_golden.py`` executes the reference's own functions out of
/root/reference and stores their outputs in ``tests/golden/glue_golden.json``;
``tests/test_oracle.py`` checks every function here against those vectors.

Functions and the reference lines they follow:
  * ``unpack_stream``      vllm_inference/modal_audio_stream.py:153-188 (canonical helper:
                           offsets, then clamp a level only if it is out of range)
  * ``unpack_trt``         tensorrt_tts/inference.py:51-93 (per-code max(0,min(4095,.)))
  * ``unpack_canopy``      tensorrt_tts/hindi_canopy/inference.py:47-60 + :171-193
  * ``token_to_code``      vllm_inference/modal_audio_stream.py:103,366
  * ``convert_to_audio``   vllm_inference/modal_audio_stream.py:132-202
  * ``decode_snac``        tensorrt_tts/inference.py:96-112
  * ``stream_chunks``      vllm_inference/modal_audio_stream.py:352-396 (buffer policy)
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

TOKEN_AUDIO_BASE = 128266       # modal_audio_stream.py:103
FRAME = 7                       # tokens per frame
WINDOW = 28                     # MIN_FRAMES_SUBSEQ, modal_audio_stream.py:92
AUDIO_SLICE_START = 2048        # modal_audio_stream.py:94
AUDIO_SLICE_END = 4096          # modal_audio_stream.py:95
POSITION_OFFSETS = [0, 4096, 8192, 12288, 16384, 20480, 24576]  # tensorrt_tts/inference.py:51


def token_to_code(token_id: int) -> int:
    return token_id - TOKEN_AUDIO_BASE


def _split(codes: Sequence[int]):
    """Frame split shared by every variant: (l0, l1, l2) before any clamp."""
    n = len(codes) // FRAME
    l0, l1, l2 = [], [], []
    for i in range(n):
        b = i * FRAME
        l0.append(codes[b])
        l1.append(codes[b + 1] - 4096)
        l2.append(codes[b + 2] - 2 * 4096)
        l2.append(codes[b + 3] - 3 * 4096)
        l1.append(codes[b + 4] - 4 * 4096)
        l2.append(codes[b + 5] - 5 * 4096)
        l2.append(codes[b + 6] - 6 * 4096)
    return l0, l1, l2


def unpack_stream(codes: Sequence[int]) -> Optional[Tuple[List[int], List[int], List[int]]]:
    """modal_audio_stream.py:153-188.  None for <7 codes; a level is clamped to
    [0, 4095] only when its min/max leave the range (equivalent to always clamping)."""
    if len(codes) < FRAME:
        return None
    levels = _split(codes)
    out = []
    for lv in levels:
        if min(lv) < 0 or max(lv) >= 4096:
            lv = [min(4095, max(0, c)) for c in lv]
        out.append(lv)
    return tuple(out)


def unpack_trt(codes: Sequence[int]) -> Tuple[List[int], List[int], List[int]]:
    """tensorrt_tts/inference.py:54-93: per-code clamp, no length guard."""
    l0, l1, l2 = _split(codes)
    cl = lambda lv: [max(0, min(4095, c)) for c in lv]
    return cl(l0), cl(l1), cl(l2)


def unpack_canopy(codes: Sequence[int]) -> Tuple[List[int], List[int], List[int]]:
    """hindi_canopy/inference.py:171-193: caller truncates to whole frames, splits with no
    clamp, then clamps each level tensor that is out of range."""
    n = len(codes) // FRAME
    l0, l1, l2 = _split(list(codes)[: n * FRAME])
    cl = lambda lv: [max(0, min(4095, c)) for c in lv] if lv and (min(lv) < 0 or max(lv) >= 4096) else lv
    return cl(l0), cl(l1), cl(l2)


def unpack_np(codes: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Vectorised form over a batch: codes int [B, 7F] (already minus 128266) ->
    int32 [B,F], [B,2F], [B,4F], clamped.  Equal to every variant above after clamping."""
    codes = np.asarray(codes, dtype=np.int64)
    B, n = codes.shape
    F_ = n // FRAME
    c = codes[:, : F_ * FRAME].reshape(B, F_, FRAME) - np.asarray(POSITION_OFFSETS, dtype=np.int64)
    c = np.clip(c, 0, 4095).astype(np.int32)
    l0 = c[:, :, 0]
    l1 = c[:, :, [1, 4]].reshape(B, 2 * F_)
    l2 = c[:, :, [2, 3, 5, 6]].reshape(B, 4 * F_)
    return l0, l1, l2


def pcm16_torch(x):
    """modal_audio_stream.py:201: (x*32767).clamp(-32768,32767).to(int16) -- truncation."""
    import torch
    return (x * 32767.0).clamp(-32768, 32767).to(torch.int16)


def pcm16_numpy(x: np.ndarray) -> np.ndarray:
    """tensorrt_tts/inference.py:108-110: clip(x,-1,1)*32767 -> astype(int16) -- truncation."""
    return (np.clip(x, -1.0, 1.0) * 32767).astype(np.int16)


def convert_to_audio(model, code_list: Sequence[int], extract_slice: bool = False, noises=None) -> Optional[bytes]:
    """modal_audio_stream.py:132-202 with ``model`` = oracle ``SnacDecodeRef``."""
    import torch
    lv = unpack_stream(list(code_list))
    if lv is None:
        return None
    codes = [torch.tensor(l, dtype=torch.int32).unsqueeze(0) for l in lv]
    audio_hat = model.decode(codes, noises)
    if extract_slice and audio_hat.shape[-1] > AUDIO_SLICE_END:
        audio_hat = audio_hat[:, :, AUDIO_SLICE_START:AUDIO_SLICE_END]
    return pcm16_torch(audio_hat).flatten().cpu().numpy().tobytes()


def decode_snac(model, layer0, layer1, layer2, noises=None) -> bytes:
    """tensorrt_tts/inference.py:96-112 with ``model`` = oracle ``SnacDecodeRef``."""
    import torch
    codes = [torch.tensor(l, dtype=torch.int32).unsqueeze(0) for l in (layer0, layer1, layer2)]
    audio = model.decode(codes, noises)
    return pcm16_numpy(audio.squeeze().cpu().numpy()).tobytes()


def stream_chunks(codes: Sequence[int]) -> List[List[int]]:
    """Buffer policy of ``stream_audio`` (modal_audio_stream.py:352-396): every time 28 codes
    are buffered they are popped and decoded; at end of stream the remaining whole frames
    (if at least one) are decoded.  Returns the list of code chunks handed to the helper."""
    out, buf = [], []
    for c in codes:
        buf.append(c)
        if len(buf) >= WINDOW:
            out.append(buf[:WINDOW])
            buf = buf[WINDOW:]
    if len(buf) >= FRAME:
        n = len(buf) // FRAME
        out.append(buf[: n * FRAME])
    return out


def sliding_windows(codes: Sequence[int]) -> List[List[int]]:
    """Sliding policy the constants describe (modal_audio_stream.py:86-95: buffer the last 28
    tokens, decode every 7 new tokens, keep samples [2048:4096]) -- the upstream Orpheus
    ``tokens_decoder`` rule the helper's ``extract_slice=True`` branch exists for."""
    out = []
    n = len(codes)
    for count in range(1, n + 1):
        if count % FRAME == 0 and count > WINDOW - 1:
            out.append(list(codes[count - WINDOW: count]))
    return out
