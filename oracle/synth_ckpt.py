"""Oracle-side checkpoint helpers: the synthetic checkpoint as an oracle model, and the one-off
gain calibration.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  The seeded generator itself
(``make_state_dict``, ``make_tokens``, ``make_noises`` ...) is product-side
(``tts_inference_b200/synth.py``: the benchmark needs random-init weights without importing
the oracle) and is re-exported here.
"""
from __future__ import annotations

import json
from typing import Dict, Optional

import numpy as np

from tts_inference_b200.synth import (  # noqa: F401
    _GAINS_PATH, load_gains, make_codes, make_noises, make_state_dict, make_tokens, noise_lengths,
    rng_bits, rng_normal, rng_uniform,
)


def make_model(seed: int = 0, gains: Optional[Dict[str, float]] = None, state_dict=None):
    """Oracle ``SnacDecodeRef`` carrying the synthetic checkpoint (fp32, eval)."""
    import torch
    from .snac_ref import SnacDecodeRef
    m = SnacDecodeRef().eval()
    if state_dict is None:
        state_dict = make_state_dict(seed, gains)
    sd = {k: torch.from_numpy(np.ascontiguousarray(v).copy()) for k, v in state_dict.items()}
    m.load_snac_state_dict(sd)
    return m


# ----------------------------------------------------------------------------------------
# one-off calibration of the per-layer gains (writes tts_inference_b200/synth_gains.json)
# ----------------------------------------------------------------------------------------

def calibrate(seed: int = 0, batch: int = 4, frames: int = 4, verbose: bool = True) -> Dict[str, float]:
    import torch
    from . import glue_ref
    targets = {"out_proj": 0.58, "dwstem": 1.0, "pwstem": 1.0, "convt": 1.0, "noise": 0.15,
               "res": 0.35, "tail": 0.5}
    gains: Dict[str, float] = {}
    l0, l1, l2 = glue_ref.unpack_np(make_codes(batch, frames))
    codes = [torch.from_numpy(x.astype(np.int64)) for x in (l0, l1, l2)]
    noises = [torch.from_numpy(n) for n in make_noises(batch, 4 * frames)]

    def run():
        m = make_model(seed, gains)
        taps: dict = {}
        m.decode(codes, noises, taps)
        return m, taps

    def fix(prefix, measured, target):
        gains[prefix] = gains.get(prefix, 1.0) * target / float(measured.detach())
        if verbose:
            print(f"{prefix:45s} std {float(measured.detach()):8.4f} -> gain {gains[prefix]:.4f}")

    m, taps = run()
    for i in range(3):
        q = m.quantizer.quantizers[i]
        z = q.out_proj(q.decode_code(codes[i]))
        fix(f"quantizer.quantizers.{i}.out_proj", z.std(), targets["out_proj"])
    m, taps = run(); fix("decoder.model.0", taps["model.0"].std(), targets["dwstem"])
    m, taps = run(); fix("decoder.model.1", taps["model.1"].std(), targets["pwstem"])
    for bi in range(4):
        p = f"decoder.model.{2 + bi}.block"
        t = f"model.{2 + bi}"
        m, taps = run(); fix(f"{p}.1", taps[f"{t}.1"].std(), targets["convt"])
        m, taps = run()
        fix(f"{p}.2.linear", ((taps[f"{t}.2"] - taps[f"{t}.1"]) / noises[bi]).std(), targets["noise"])
        for ri in range(3):
            m, taps = run()
            fix(f"{p}.{3 + ri}.block.3", (taps[f"{t}.{3 + ri}"] - taps[f"{t}.{2 + ri}"]).std(), targets["res"])
    m, taps = run(); fix("decoder.model.7", taps["model.7"].std(), targets["tail"])
    with open(_GAINS_PATH, "w") as f:
        json.dump({k: round(v, 6) for k, v in gains.items()}, f, indent=1, sort_keys=True)
    m, taps = run()
    if verbose:
        for k, v in taps.items():
            print(f"{k:14s} shape {tuple(v.shape)} mean {float(v.mean()):8.4f} std {float(v.std()):8.4f} "
                  f"absmax {float(v.abs().max()):8.3f}")
    return gains


if __name__ == "__main__":
    calibrate()
