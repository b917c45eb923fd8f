"""Shared helpers of the parity tests (test infrastructure; may import the oracle)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from oracle import glue_ref

TOKEN_AUDIO_BASE = 128266


def oracle_decode(model, tokens: np.ndarray, noises: Optional[List[np.ndarray]], raw_ids: bool = True,
                  want_taps: bool = False) -> Tuple[np.ndarray, Dict[str, np.ndarray]]:
    """tokens int [B, 7F] -> (wave float32 [B, 2048F], taps keyed like snacb_debug_tap, channel-last)."""
    codes = tokens.astype(np.int64) - (TOKEN_AUDIO_BASE if raw_ids else 0)
    lv = glue_ref.unpack_np(codes)
    taps: dict = {} if want_taps else None
    nz = None if noises is None else [torch.from_numpy(np.ascontiguousarray(n)) for n in noises]
    y = model.decode([torch.from_numpy(x.astype(np.int64)) for x in lv], nz, taps)
    out: Dict[str, np.ndarray] = {}
    if want_taps:
        def cl(t):   # [B,C,T] -> [B*T, C]
            return t.permute(0, 2, 1).reshape(-1, t.shape[1]).contiguous().numpy()
        out["stem_dw"] = cl(taps["model.0"])
        out["stem"] = cl(taps["model.2.0"])
        for i in range(4):
            p = f"model.{2 + i}"
            out[f"b{i}.convt"] = cl(taps[f"{p}.1"])
            out[f"b{i}.noise"] = cl(taps[f"{p}.2"])
            out[f"b{i}.res0"] = cl(taps[f"{p}.3"])
            out[f"b{i}.res1"] = cl(taps[f"{p}.4"])
            out[f"b{i}.res2"] = cl(taps[f"model.{3 + i}.0"] if i < 3 else taps["model.6"])
    return y[:, 0].numpy(), out


def snr_db(ref: np.ndarray, got: np.ndarray) -> float:
    ref = ref.astype(np.float64); got = got.astype(np.float64)
    err = ((ref - got) ** 2).sum()
    return float(10 * np.log10((ref ** 2).sum() / max(err, 1e-300)))


def pcm_of(wave: np.ndarray) -> np.ndarray:
    return glue_ref.pcm16_torch(torch.from_numpy(wave)).numpy()
