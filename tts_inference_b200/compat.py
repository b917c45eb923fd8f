"""Drop-in mirror of the reference's SNAC helper -- same names, arguments and error behaviour.

    from tts_inference_b200.compat import init_snac, convert_to_audio      # vLLM streaming path
    from tts_inference_b200.compat import redistribute_codes, decode_snac  # TensorRT batch path

Reference                                            here
---------------------------------------------------  -------------------------------------------
init_snac()            modal_audio_stream.py:106-129  init_snac(checkpoint=None, device=0)   [None -> $SNACB_CKPT, else raises]
convert_to_audio()     modal_audio_stream.py:132-202  convert_to_audio(code_list, extract_slice=False)
redistribute_codes()   tensorrt_tts/inference.py:54-93   redistribute_codes(codes)
decode_snac()          tensorrt_tts/inference.py:96-112  decode_snac(l0, l1, l2, snac_model, device)

``snac_model`` / ``snac_device`` are module globals exactly as in the reference
(modal_audio_stream.py:79-80).  Everything numeric runs in the CUDA library; there is no CPU path.
"""
from __future__ import annotations

import itertools
from typing import List, Optional, Tuple

import numpy as np

from .api import FRAME, SnacDecoder

snac_model: Optional[SnacDecoder] = None
snac_device: Optional[str] = None

AUDIO_SLICE_START = 2048     # modal_audio_stream.py:94
AUDIO_SLICE_END = 4096       # modal_audio_stream.py:95
TOKEN_AUDIO_BASE = 128266    # modal_audio_stream.py:103
POSITION_OFFSETS = [0, 4096, 8192, 12288, 16384, 20480, 24576]   # tensorrt_tts/inference.py:51

_noise_seed = itertools.count(1)      # the reference draws fresh torch.randn noise on every decode
_I32_MIN, _I32_MAX = -(2 ** 31), 2 ** 31 - 1


# Test-only hooks (tests/test_gpu_api.py): the reference helper draws fresh torch.randn noise inside snac.decode and
# runs fp32; to compare the bytes these entry points return with golden bytes of the reference's own functions, a test
# pins the NoiseBlock noise (list of four cuda float32 tensors [1, 1, T_i]) and the arithmetic ("fp32").
_test_noise = None
_test_precision = None


def init_snac(checkpoint=None, device: int = 0) -> SnacDecoder:
    """Load SNAC weights onto the GPU (weight-norm folded, tensor-core tiles packed).

    The reference takes no argument and downloads ``hubertsiuzdak/snac_24khz``
    (modal_audio_stream.py:113).  There is no network here, so the checkpoint is
    ``checkpoint`` -- path of ``pytorch_model.bin`` / its directory, or a state dict -- or, for the
    reference's argument-less call, the path in the environment variable ``SNACB_CKPT``.  With neither
    this raises: weights are never made up.  (Benchmarks and tests that want random-init weights pass
    ``synth.make_state_dict(seed)`` explicitly.)  The reference's warm-up decode
    (modal_audio_stream.py:120-127) is kept: it also sizes the workspace."""
    global snac_model, snac_device
    import os
    from . import weights
    if checkpoint is None:
        checkpoint = os.environ.get(weights.CKPT_ENV)
        if not checkpoint:
            raise RuntimeError(
                "init_snac(): no checkpoint.  Pass the path of snac_24khz's pytorch_model.bin (or its directory, or a "
                f"state dict), or set {weights.CKPT_ENV}; this build cannot download hubertsiuzdak/snac_24khz.")
    if isinstance(checkpoint, (str, os.PathLike)):
        model = SnacDecoder(weights.load_folded(os.fspath(checkpoint)), device=device, folded=True)
    else:
        model = SnacDecoder(checkpoint, device=device)
    snac_model = model
    snac_device = f"cuda:{device}"
    snac_model.decode_host(np.zeros((1, FRAME), dtype=np.int32))        # warm-up, one frame
    return snac_model


def _decode_one(model: SnacDecoder, tok: np.ndarray, extract_slice: bool) -> np.ndarray:
    if _test_noise is None and _test_precision is None:
        return model.decode_host(tok, raw_ids=False, extract_slice=extract_slice, seed=next(_noise_seed))
    import torch
    t = torch.from_numpy(tok).to(f"cuda:{model.device}")
    pcm = model.decode(t, raw_ids=False, extract_slice=extract_slice, noise=_test_noise, seed=next(_noise_seed),
                       precision=_test_precision or "fp16")
    return pcm.cpu().numpy()


def _as_i32(codes) -> np.ndarray:
    # Python ints can exceed int32; values that far out clamp to 0/4095 either way
    return np.asarray([min(_I32_MAX, max(_I32_MIN, int(c))) for c in codes], dtype=np.int32)


def convert_to_audio(code_list: list, extract_slice: bool = False) -> Optional[bytes]:
    """Codes (``token_id - 128266``) of one stream -> PCM bytes (int16 LE, 24 kHz mono).
    None for fewer than 7 codes; a trailing partial frame is dropped; out-of-range codes clamp to
    [0, 4095]; ``extract_slice`` keeps samples [2048:4096] when more than 4096 were decoded."""
    if snac_model is None:
        raise RuntimeError("init_snac() has not been called")
    if len(code_list) < FRAME:
        return None
    num_frames = len(code_list) // FRAME
    tok = _as_i32(code_list[: num_frames * FRAME]).reshape(1, -1)
    return _decode_one(snac_model, tok, extract_slice).tobytes()


def redistribute_codes(codes: List[int]) -> Tuple[List[int], List[int], List[int]]:
    """Flat codes -> (layer0, layer1, layer2), offsets removed and clamped to [0, 4095]
    (integer kernel on the GPU, bit-exact with the reference loop)."""
    if snac_model is None:
        raise RuntimeError("init_snac() has not been called")
    import torch
    num_frames = len(codes) // FRAME
    if num_frames == 0:
        return [], [], []
    tok = torch.from_numpy(_as_i32(codes[: num_frames * FRAME]).reshape(1, -1)).to(snac_device)
    c0, c1, c2 = snac_model.unpack(tok, raw_ids=False)
    return c0[0].tolist(), c1[0].tolist(), c2[0].tolist()


def decode_snac(layer0: List[int], layer1: List[int], layer2: List[int], snac_model_=None, device: str = "cuda") -> bytes:
    """Three code levels -> PCM bytes.  The levels are re-interleaved into frame order and go
    through the same fused decode (``snac_model_`` / ``device`` are accepted for signature parity)."""
    model = snac_model_ if isinstance(snac_model_, SnacDecoder) else snac_model
    if model is None:
        raise RuntimeError("init_snac() has not been called")
    F_ = len(layer0)
    if F_ == 0:
        return b""
    l0 = np.asarray(layer0, dtype=np.int64)
    l1 = np.asarray(layer1, dtype=np.int64).reshape(F_, 2)
    l2 = np.asarray(layer2, dtype=np.int64).reshape(F_, 4)
    flat = np.stack([l0, l1[:, 0], l2[:, 0], l2[:, 1], l1[:, 1], l2[:, 2], l2[:, 3]], axis=1)
    flat = np.clip(flat, 0, 4095) + np.asarray(POSITION_OFFSETS, dtype=np.int64)[None, :]
    return _decode_one(model, np.ascontiguousarray(flat.reshape(1, -1).astype(np.int32)), False).tobytes()
