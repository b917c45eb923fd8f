"""Checkpoint -> folded fp32 weights for ``snacb_create`` (host-side load logic).

The reference keeps weight-norm un-folded and recomputes ``w = g * v / ||v||`` inside every
forward of the pip ``snac`` model it loads in ``init_snac`` (vllm_inference/modal_audio_stream.py:
106-129).  Here the fold happens once at load.  Norm is over every dim but 0 (torch
``weight_norm(dim=0)``): per OUTPUT channel for Conv1d ``[Cout, Cin/groups, k]``, per INPUT
channel for ConvTranspose1d ``[Cin, Cout, k]``.  Both key styles are accepted:
``*.weight_g`` / ``*.weight_v`` and ``*.parametrizations.weight.original0`` / ``original1``.
"""
from __future__ import annotations

from typing import Dict, Mapping

import numpy as np

DECODER_RATES = (8, 8, 4, 2)


def _np(x) -> np.ndarray:
    if isinstance(x, np.ndarray):
        return x
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def fold_weight_norm(g, v) -> np.ndarray:
    """w = g * v / ||v||, norm over all dims except 0; float64 arithmetic, float32 result."""
    v64 = _np(v).astype(np.float64)
    g64 = _np(g).astype(np.float64).reshape(v64.shape[0], *([1] * (v64.ndim - 1)))
    norm = np.sqrt((v64.reshape(v64.shape[0], -1) ** 2).sum(axis=1)).reshape(g64.shape)
    return np.ascontiguousarray((v64 * (g64 / norm)).astype(np.float32))


def _weight(sd: Mapping[str, object], prefix: str) -> np.ndarray:
    if prefix + ".weight_g" in sd:
        return fold_weight_norm(sd[prefix + ".weight_g"], sd[prefix + ".weight_v"])
    p = prefix + ".parametrizations.weight.original"
    if p + "0" in sd:
        return fold_weight_norm(sd[p + "0"], sd[p + "1"])
    if prefix + ".weight" in sd:        # already folded checkpoint
        return np.ascontiguousarray(_np(sd[prefix + ".weight"]).astype(np.float32))
    raise KeyError(f"checkpoint has no weight for {prefix}")


def _vec(sd, key) -> np.ndarray:
    if key not in sd:
        raise KeyError(f"checkpoint lacks {key}")
    return np.ascontiguousarray(_np(sd[key]).astype(np.float32).reshape(-1))


def fold_state_dict(sd: Mapping[str, object]) -> Dict[str, np.ndarray]:
    """Flat dict of contiguous float32 arrays, keyed by the field names of ``snacb_weights``."""
    out: Dict[str, np.ndarray] = {}
    for i in range(3):
        q = f"quantizer.quantizers.{i}"
        out[f"codebook{i}"] = np.ascontiguousarray(_np(sd[q + ".codebook.weight"]).astype(np.float32))
        out[f"out_proj_w{i}"] = _weight(sd, q + ".out_proj")
        out[f"out_proj_b{i}"] = _vec(sd, q + ".out_proj.bias")
        assert out[f"codebook{i}"].shape == (4096, 8) and out[f"out_proj_w{i}"].shape == (768, 8, 1)
    out["stem_dw_w"] = _weight(sd, "decoder.model.0")
    out["stem_dw_b"] = _vec(sd, "decoder.model.0.bias")
    out["stem_pw_w"] = _weight(sd, "decoder.model.1")
    out["stem_pw_b"] = _vec(sd, "decoder.model.1.bias")
    assert out["stem_dw_w"].shape == (768, 1, 7) and out["stem_pw_w"].shape == (1024, 768, 1)
    cin = 1024
    for bi, s in enumerate(DECODER_RATES):
        cout = cin // 2
        p = f"decoder.model.{2 + bi}.block"
        out[f"b{bi}.alpha"] = _vec(sd, p + ".0.alpha")
        out[f"b{bi}.convt_w"] = _weight(sd, p + ".1")
        out[f"b{bi}.convt_b"] = _vec(sd, p + ".1.bias")
        out[f"b{bi}.noise_w"] = _weight(sd, p + ".2.linear")
        assert out[f"b{bi}.convt_w"].shape == (cin, cout, 2 * s), out[f"b{bi}.convt_w"].shape
        assert out[f"b{bi}.noise_w"].shape == (cout, cout, 1)
        for ri in range(3):
            q = f"{p}.{3 + ri}.block"
            out[f"b{bi}.r{ri}.alpha1"] = _vec(sd, q + ".0.alpha")
            out[f"b{bi}.r{ri}.dw_w"] = _weight(sd, q + ".1")
            out[f"b{bi}.r{ri}.dw_b"] = _vec(sd, q + ".1.bias")
            out[f"b{bi}.r{ri}.alpha2"] = _vec(sd, q + ".2.alpha")
            out[f"b{bi}.r{ri}.pw_w"] = _weight(sd, q + ".3")
            out[f"b{bi}.r{ri}.pw_b"] = _vec(sd, q + ".3.bias")
            assert out[f"b{bi}.r{ri}.dw_w"].shape == (cout, 1, 7) and out[f"b{bi}.r{ri}.pw_w"].shape == (cout, cout, 1)
        cin = cout
    out["tail_alpha"] = _vec(sd, "decoder.model.6.alpha")
    out["tail_w"] = _weight(sd, "decoder.model.7")
    out["tail_b"] = _vec(sd, "decoder.model.7.bias")
    assert out["tail_w"].shape == (1, 64, 7)
    return out


ENCODER_RATES = (2, 4, 8, 8)


def fold_encoder_state_dict(sd: Mapping[str, object]) -> Dict[str, np.ndarray]:
    """Encode half of the checkpoint (``encoder.block.*`` and the quantizers' in_proj / codebook / out_proj), folded, keyed
    by the field names of ``snacb_encoder_weights`` (include/snacb.h)."""
    out: Dict[str, np.ndarray] = {}
    out["enc.conv0_w"] = _weight(sd, "encoder.block.0")
    out["enc.conv0_b"] = _vec(sd, "encoder.block.0.bias")
    assert out["enc.conv0_w"].shape == (48, 1, 7)
    c = 48
    for bi, s in enumerate(ENCODER_RATES):
        p = f"encoder.block.{1 + bi}.block"
        for ri in range(3):
            q = f"{p}.{ri}.block"
            out[f"enc.b{bi}.r{ri}.alpha1"] = _vec(sd, q + ".0.alpha")
            out[f"enc.b{bi}.r{ri}.dw_w"] = _weight(sd, q + ".1")
            out[f"enc.b{bi}.r{ri}.dw_b"] = _vec(sd, q + ".1.bias")
            out[f"enc.b{bi}.r{ri}.alpha2"] = _vec(sd, q + ".2.alpha")
            out[f"enc.b{bi}.r{ri}.pw_w"] = _weight(sd, q + ".3")
            out[f"enc.b{bi}.r{ri}.pw_b"] = _vec(sd, q + ".3.bias")
            assert out[f"enc.b{bi}.r{ri}.dw_w"].shape == (c, 1, 7) and out[f"enc.b{bi}.r{ri}.pw_w"].shape == (c, c, 1)
        out[f"enc.b{bi}.alpha"] = _vec(sd, p + ".3.alpha")
        out[f"enc.b{bi}.conv_w"] = _weight(sd, p + ".4")
        out[f"enc.b{bi}.conv_b"] = _vec(sd, p + ".4.bias")
        assert out[f"enc.b{bi}.conv_w"].shape == (2 * c, c, 2 * s), out[f"enc.b{bi}.conv_w"].shape
        c *= 2
    out["enc.final_w"] = _weight(sd, "encoder.block.5")
    out["enc.final_b"] = _vec(sd, "encoder.block.5.bias")
    assert out["enc.final_w"].shape == (768, 1, 7)
    for i in range(3):
        q = f"quantizer.quantizers.{i}"
        out[f"in_proj_w{i}"] = _weight(sd, q + ".in_proj")
        out[f"in_proj_b{i}"] = _vec(sd, q + ".in_proj.bias")
        out[f"codebook{i}"] = np.ascontiguousarray(_np(sd[q + ".codebook.weight"]).astype(np.float32))
        out[f"out_proj_w{i}"] = _weight(sd, q + ".out_proj")
        out[f"out_proj_b{i}"] = _vec(sd, q + ".out_proj.bias")
        assert out[f"in_proj_w{i}"].shape == (8, 768, 1) and out[f"codebook{i}"].shape == (4096, 8)
    return out


CKPT_ENV = "SNACB_CKPT"            # path of pytorch_model.bin (or its directory) used when init_snac() gets no argument
CACHE_ENV = "SNACB_CACHE_DIR"      # where folded weights are cached (default ~/.cache/snacb); "" disables the cache
_FOLD_VERSION = 1                  # bump when fold_state_dict's output changes


def resolve_checkpoint(path: str) -> str:
    """``path`` may be the file or the directory ``SNAC.from_pretrained`` would have downloaded into
    (``pytorch_model.bin`` next to ``config.json``)."""
    import os
    if os.path.isdir(path):
        path = os.path.join(path, "pytorch_model.bin")
    if not os.path.isfile(path):
        raise FileNotFoundError(f"SNAC checkpoint not found: {path}")
    return path


def load_checkpoint(path: str) -> Dict[str, np.ndarray]:
    """Read ``pytorch_model.bin`` (or a directory holding it) of hubertsiuzdak/snac_24khz: the raw state dict
    (weight-norm NOT folded), either key style."""
    import torch
    sd = torch.load(resolve_checkpoint(path), map_location="cpu", weights_only=True)
    if isinstance(sd, dict) and "state_dict" in sd and "decoder.model.0.bias" not in sd:
        sd = sd["state_dict"]
    return {k: v.numpy() for k, v in sd.items() if hasattr(v, "numpy")}


def _file_sha256(path: str) -> str:
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 22), b""):
            h.update(chunk)
    return h.hexdigest()


def load_folded(path: str, cache_dir: "str | None" = None) -> Dict[str, np.ndarray]:
    """Checkpoint file -> folded fp32 arrays (``fold_state_dict``), through a cache keyed by the file's SHA-256:
    the reference re-derives ``g * v / ||v||`` on every forward; here it is derived once per checkpoint CONTENT and
    later loads of the same file skip torch.load and the fold.  A stale or unreadable cache entry
    is ignored and rewritten."""
    import os
    path = resolve_checkpoint(path)
    if cache_dir is None:
        cache_dir = os.environ.get(CACHE_ENV, os.path.join(os.path.expanduser("~"), ".cache", "snacb"))
    entry = None
    if cache_dir:
        entry = os.path.join(cache_dir, f"folded-v{_FOLD_VERSION}-{_file_sha256(path)[:32]}.npz")
        if os.path.isfile(entry):
            try:
                with np.load(entry) as z:
                    out = {k: np.ascontiguousarray(z[k], dtype=np.float32) for k in z.files}
                if "tail_w" in out and "b3.r2.pw_w" in out:
                    return out
            except Exception:  # noqa: BLE001 -- corrupt entry: fall through and rebuild it
                pass
    out = fold_state_dict(load_checkpoint(path))
    if entry:
        try:
            os.makedirs(cache_dir, exist_ok=True)
            tmp = f"{entry}.{os.getpid()}.tmp.npz"
            np.savez(tmp, **out)
            os.replace(tmp, entry)
        except OSError:
            pass                     # read-only home etc.: the cache is an optimisation only
    return out
