"""Deterministic synthetic SNAC-24k checkpoint, token windows and noise.

Random-init weights of the snac_24khz decode architecture plus synthetic inputs, for the
benchmark, the smoke test and the parity tests.  The real checkpoint
(``hubertsiuzdak/snac_24khz``, fetched by the reference at
vllm_inference/modal_audio_stream.py:113) cannot be downloaded here, so everything runs on
a seeded synthetic one with the same tensor names/shapes.  Values come from a counter-based
generator (splitmix64 -> Box-Muller in float64) written out below so that they do not
depend on any library's RNG stream; per-layer gains were calibrated once
(``oracle/synth_ckpt.py::calibrate``)
so that activations stay O(1) through the stack and the pre-tanh signal has std ~0.5
(a saturated tanh would hide errors), and are committed in ``synth_gains.json``.
"""
from __future__ import annotations

import json
import os
import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np

_GAINS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth_gains.json")
TOKEN_AUDIO_BASE = 128266

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        z = z ^ (z >> np.uint64(31))
    return z


def rng_bits(seed: int, stream: int, n: int, offset: int = 0) -> np.ndarray:
    """n 64-bit words of stream (seed, stream): splitmix64(splitmix64(key) + counter)."""
    key = _splitmix64(np.array([(seed * 0x100000001B3 + stream) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
    ctr = np.arange(offset, offset + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return _splitmix64((ctr + key) & _M64)


def rng_uniform(seed: int, stream: int, n: int, offset: int = 0) -> np.ndarray:
    b = rng_bits(seed, stream, n, offset)
    return ((b >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def rng_normal(seed: int, stream: int, n: int, offset: int = 0) -> np.ndarray:
    b = rng_bits(seed, stream, n, offset)
    u1 = ((b >> np.uint64(32)).astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
    u2 = ((b & np.uint64(0xFFFFFFFF)).astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def _sid(name: str) -> int:
    return zlib.crc32(name.encode())


# ----------------------------------------------------------------------------------------
# checkpoint
# ----------------------------------------------------------------------------------------

def _layer_table():
    """(prefix, kind, shape_v, g_len, has_bias) for every weight-normed conv on the decode
    path, in forward order.  kind in {out_proj, dw, pw, convt, noise, tail}."""
    rows = []
    for i in range(3):
        rows.append((f"quantizer.quantizers.{i}.out_proj", "out_proj", (768, 8, 1), 768, True))
    rows.append(("decoder.model.0", "dw", (768, 1, 7), 768, True))
    rows.append(("decoder.model.1", "pw", (1024, 768, 1), 1024, True))
    cin = 1024
    for bi, s in enumerate((8, 8, 4, 2)):
        cout = cin // 2
        p = f"decoder.model.{2 + bi}.block"
        rows.append((f"{p}.1", "convt", (cin, cout, 2 * s), cin, True))
        rows.append((f"{p}.2.linear", "noise", (cout, cout, 1), cout, False))
        for ri in range(3):
            q = f"{p}.{3 + ri}.block"
            rows.append((f"{q}.1", "dw", (cout, 1, 7), cout, True))
            rows.append((f"{q}.3", "pw", (cout, cout, 1), cout, True))
        cin = cout
    rows.append(("decoder.model.7", "tail", (1, 64, 7), 1, True))
    return rows


def _alpha_keys():
    keys = []
    cin = 1024
    for bi in range(4):
        cout = cin // 2
        p = f"decoder.model.{2 + bi}.block"
        keys.append((f"{p}.0.alpha", cin))
        for ri in range(3):
            keys.append((f"{p}.{3 + ri}.block.0.alpha", cout))
            keys.append((f"{p}.{3 + ri}.block.2.alpha", cout))
        cin = cout
    keys.append(("decoder.model.6.alpha", 64))
    return keys


def load_gains() -> Dict[str, float]:
    if os.path.exists(_GAINS_PATH):
        with open(_GAINS_PATH) as f:
            return json.load(f)
    return {}


ALPHA_MODES = ("benign", "wild", "hard")


def _alphas(seed: int, key: str, c: int, mode: str) -> np.ndarray:
    """Snake alpha of one layer.  ``benign``: U(0.3, 3) (what the gains were calibrated on).  Trained alphas can be
    ~0, negative or large, so the parity tests also run on
      ``wild``: half of the channels from {-0.7, 12, 40} -- still inside the range where the fp16 chain kernel's
                alpha-folded formulation is used (csrc/snacb.cu chain_fold_safe), and
      ``hard``: half of the channels from {0, 1e-4, -1e-4, -0.7, 12, 40} -- forces the general (fp32 Snake) variant;
                alpha = 0 must give snake(x) = x as in the reference."""
    base = 0.3 + 2.7 * rng_uniform(seed, _sid(key), c)
    if mode == "benign":
        return base
    pool = np.array([-0.7, 12.0, 40.0] if mode == "wild" else [0.0, 1e-4, -1e-4, -0.7, 12.0, 40.0])
    pick = rng_uniform(seed, _sid(key + ".pick"), c) < 0.5
    which = (rng_bits(seed, _sid(key + ".which"), c) % np.uint64(len(pool))).astype(np.int64)
    return np.where(pick, pool[which], base)


def make_state_dict(seed: int = 0, gains: Optional[Dict[str, float]] = None, alpha_mode: str = "benign",
                    act_scale: float = 1.0) -> Dict[str, "np.ndarray"]:
    """Synthetic decode-path state dict as float32 numpy arrays, upstream key names
    (old-style ``weight_g``/``weight_v``).  g = gain * ||v|| * U(0.5, 1.5) per norm channel.
    ``act_scale``: activations of the whole decoder ``act_scale`` times larger (stem 1x1 gain up, tail conv gain down
    by the same factor, so the tanh still is not saturated)."""
    assert alpha_mode in ALPHA_MODES
    if gains is None:
        gains = load_gains()
    if act_scale != 1.0:
        gains = dict(gains)
        gains["decoder.model.1"] = gains.get("decoder.model.1", 1.0) * act_scale
        gains["decoder.model.7"] = gains.get("decoder.model.7", 1.0) / act_scale
    sd: Dict[str, np.ndarray] = {}
    for i in range(3):
        k = f"quantizer.quantizers.{i}.codebook.weight"
        sd[k] = rng_normal(seed, _sid(k), 4096 * 8).reshape(4096, 8).astype(np.float32)
    for prefix, kind, shape, glen, has_bias in _layer_table():
        n = int(np.prod(shape))
        v = rng_normal(seed, _sid(prefix + ".v"), n).reshape(shape)
        fan_in = {"out_proj": 8, "dw": 7, "pw": shape[1], "noise": shape[1],
                  "convt": 2 * shape[0], "tail": 64 * 7}[kind]
        v = v / np.sqrt(fan_in)
        norm = np.sqrt((v.reshape(shape[0], -1) ** 2).sum(axis=1))          # dim=0 norm
        jitter = 0.5 + rng_uniform(seed, _sid(prefix + ".g"), glen)
        gain = float(gains.get(prefix, 1.0))
        g = (gain * norm * jitter).reshape(shape[0], 1, 1)
        sd[prefix + ".weight_v"] = v.astype(np.float32)
        sd[prefix + ".weight_g"] = g.astype(np.float32)
        if has_bias:
            nb = shape[1] if kind == "convt" else shape[0]
            sd[prefix + ".bias"] = (0.1 * rng_normal(seed, _sid(prefix + ".b"), nb)).astype(np.float32)
    for k, c in _alpha_keys():
        sd[k] = _alphas(seed, k, c, alpha_mode).reshape(1, c, 1).astype(np.float32)
    return sd


def _encoder_layer_table():
    """(prefix, kind, shape_v, gain) of every weight-normed conv of the ENCODE half, forward order."""
    rows = [("encoder.block.0", "conv0", (48, 1, 7), 3.0)]
    c = 48
    for bi, s in enumerate((2, 4, 8, 8)):
        p = f"encoder.block.{1 + bi}.block"
        for ri in range(3):
            rows.append((f"{p}.{ri}.block.1", "dw", (c, 1, 7), 1.0))
            rows.append((f"{p}.{ri}.block.3", "pw", (c, c, 1), 0.3))
        rows.append((f"{p}.4", "down", (2 * c, c, 2 * s), 0.8))
        c *= 2
    rows.append(("encoder.block.5", "dw", (768, 1, 7), 1.0))
    for i in range(3):
        rows.append((f"quantizer.quantizers.{i}.in_proj", "pw", (8, 768, 1), 1.0))
    return rows


def make_encoder_state_dict(seed: int = 0) -> Dict[str, "np.ndarray"]:
    """Synthetic ENCODE-half state dict (upstream key names: ``encoder.block.*``, ``quantizer.quantizers.i.in_proj``) on top
    of the decode checkpoint of the same seed, whose codebooks and out_proj the quantizer shares: one dict that
    ``weights.fold_encoder_state_dict`` and ``weights.fold_state_dict`` both accept.  Fan-in scaled weights with fixed
    per-kind gains; activations stay O(1) through the stack (oracle/snac_enc_ref.py, tests/test_oracle.py)."""
    sd = make_state_dict(seed)
    for prefix, kind, shape, gain in _encoder_layer_table():
        n = int(np.prod(shape))
        v = rng_normal(seed, _sid(prefix + ".v"), n).reshape(shape) / np.sqrt(shape[1] * shape[2])
        norm = np.sqrt((v.reshape(shape[0], -1) ** 2).sum(axis=1))
        jitter = 0.5 + rng_uniform(seed, _sid(prefix + ".g"), shape[0])
        sd[prefix + ".weight_v"] = v.astype(np.float32)
        sd[prefix + ".weight_g"] = (gain * norm * jitter).reshape(shape[0], 1, 1).astype(np.float32)
        sd[prefix + ".bias"] = (0.1 * rng_normal(seed, _sid(prefix + ".b"), shape[0])).astype(np.float32)
    c = 48
    for bi in range(4):
        p = f"encoder.block.{1 + bi}.block"
        for ri in range(3):
            for j in (0, 2):
                k = f"{p}.{ri}.block.{j}.alpha"
                sd[k] = _alphas(seed, k, c, "benign").reshape(1, c, 1).astype(np.float32)
        k = f"{p}.3.alpha"
        sd[k] = _alphas(seed, k, c, "benign").reshape(1, c, 1).astype(np.float32)
        c *= 2
    return sd


def make_audio(batch: int, n_samples: int, seed: int = 5) -> np.ndarray:
    """Synthetic audio float32 [B, n] in [-1, 1]: a few decaying sinusoids plus noise per row."""
    t = np.arange(n_samples, dtype=np.float64) / 24000.0
    out = np.empty((batch, n_samples), dtype=np.float32)
    for b in range(batch):
        u = rng_uniform(seed, 900 + b, 12)
        x = 0.05 * rng_normal(seed, 1000 + b, n_samples)
        for k in range(4):
            f0 = 80.0 * (1.0 + 40.0 * u[3 * k]) 
            x = x + (0.1 + 0.3 * u[3 * k + 1]) * np.sin(2 * np.pi * f0 * t + 6.28 * u[3 * k + 2])
        out[b] = np.clip(x * 0.5, -1.0, 1.0).astype(np.float32)
    return out


# ----------------------------------------------------------------------------------------
# inputs
# ----------------------------------------------------------------------------------------

def make_codes(batch: int, frames: int, seed: int = 20241224) -> np.ndarray:
    """Flat codes (token_id - 128266) [B, 7F]: c ~ U{0..4095} + 4096*(p mod 7)."""
    n = batch * frames * 7
    c = (rng_bits(seed, 1, n) % np.uint64(4096)).astype(np.int64).reshape(batch, frames * 7)
    return c + 4096 * (np.arange(frames * 7, dtype=np.int64) % 7)[None, :]


def make_tokens(batch: int, frames: int, seed: int = 20241224, bad_frac: float = 0.0) -> np.ndarray:
    """Raw LLM token ids int32 [B, 7F].  ``bad_frac`` of them are replaced by out-of-range
    ids (specials 128257..128265, ids past the audio range, wrong-position codes) to drive
    the clamp of modal_audio_stream.py:183-188 ("Hindi vocabulary" config)."""
    ids = make_codes(batch, frames, seed) + TOKEN_AUDIO_BASE
    if bad_frac > 0:
        n = ids.size
        flat = ids.reshape(-1)
        pick = rng_uniform(seed, 2, n) < bad_frac
        kind = (rng_bits(seed, 3, n) % np.uint64(3)).astype(np.int64)
        r = rng_bits(seed, 4, n)
        special = 128257 + (r % np.uint64(9)).astype(np.int64)
        beyond = TOKEN_AUDIO_BASE + 28672 + (r % np.uint64(64)).astype(np.int64)
        wrongpos = TOKEN_AUDIO_BASE + (r % np.uint64(28672)).astype(np.int64)
        bad = np.where(kind == 0, special, np.where(kind == 1, beyond, wrongpos))
        flat[pick] = bad[pick]
    return ids.astype(np.int32)


def noise_lengths(t0: int) -> List[int]:
    out, t = [], t0
    for r in (8, 8, 4, 2):
        t *= r
        out.append(t)
    return out


def make_noises_rng(batch: int, t0: int, seed: int, stream_offset: int = 0) -> List[np.ndarray]:
    """The values the kernels' built-in counter RNG draws (csrc/common.cuh noise_counter): stream key
    (seed, 100 + block), counter (stream << 32) | t -- independent of the decoded length and of the batch split."""
    out = []
    for i, t in enumerate(noise_lengths(t0)):
        rows = [rng_normal(seed, 100 + i, t, offset=(stream_offset + s) << 32) for s in range(batch)]
        out.append(np.stack(rows).reshape(batch, 1, t).astype(np.float32))
    return out


def make_noises(batch: int, t0: int, seed: int = 7) -> List[np.ndarray]:
    """Injected NoiseBlock tensors, float32 [B,1,T_i] for the four decoder blocks."""
    return [rng_normal(seed, 100 + i, batch * t).reshape(batch, 1, t).astype(np.float32)
            for i, t in enumerate(noise_lengths(t0))]
