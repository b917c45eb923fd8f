"""PyTorch fp32 restatement of ``snac.SNAC.encode`` for the 24 kHz checkpoint (ORACLE).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  PARITY UNPINNED at this boundary, for the same reason as
``snac_ref.py``: the upstream package ``snac`` (github hubertsiuzdak/snac, installed unpinned by the reference,
vllm_inference/modal_audio_stream.py:58) is not in this image, so this file restates its published module graph
(``snac/snac.py`` SNAC.preprocess / encode, ``snac/layers.py`` Encoder / EncoderBlock / ResidualUnit,
``snac/vq.py`` VectorQuantize.forward / decode_latents, ResidualVectorQuantize.forward).  The reference never calls
``encode`` at inference; what pins the structure instead: the total parameter count of encoder +
quantizer + decoder = 19.8 M (``tests/test_oracle.py``), the shared DAC lineage of
EncoderBlock / ResidualUnit / the factorised, L2-normalised codebook lookup (``transformers`` ``modeling_dac.py``), and
``from_codes(encode(z)) == z_q`` of the forward pass.

snac_24khz/config.json: encoder_dim 48, encoder_rates [2, 4, 8, 8] (hop 512), attn_window_size null (no LocalMHA),
depthwise true, codebook_size 4096, codebook_dim 8, vq_strides [4, 2, 1].
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .snac_ref import (CODEBOOK_DIM, CODEBOOK_SIZE, HOP, LATENT_DIM, VQ_STRIDES, ResidualUnit, Snake1d, VectorQuantize,
                       WNConv1d)

ENCODER_DIM = 48
ENCODER_RATES = (2, 4, 8, 8)


class EncoderBlock(nn.Module):
    """3 ResidualUnits (dilations 1, 3, 9) on the INPUT width, Snake, strided conv k = 2s, padding ceil(s / 2)."""

    def __init__(self, output_dim: int, stride: int, groups: int):
        super().__init__()
        input_dim = output_dim // 2
        self.block = nn.Sequential(
            ResidualUnit(input_dim, dilation=1, groups=groups),
            ResidualUnit(input_dim, dilation=3, groups=groups),
            ResidualUnit(input_dim, dilation=9, groups=groups),
            Snake1d(input_dim),
            WNConv1d(input_dim, output_dim, kernel_size=2 * stride, stride=stride, padding=math.ceil(stride / 2)),
        )

    def forward(self, x):
        return self.block(x)


class Encoder(nn.Module):
    def __init__(self, d_model: int = ENCODER_DIM, strides=ENCODER_RATES, depthwise: bool = True):
        super().__init__()
        layers: List[nn.Module] = [WNConv1d(1, d_model, kernel_size=7, padding=3)]
        for stride in strides:
            d_model *= 2
            groups = d_model // 2 if depthwise else 1
            layers.append(EncoderBlock(output_dim=d_model, stride=stride, groups=groups))
        groups = d_model if depthwise else 1
        layers.append(WNConv1d(d_model, d_model, kernel_size=7, padding=3, groups=groups))
        self.block = nn.Sequential(*layers)
        self.enc_dim = d_model

    def forward(self, x, taps: Optional[dict] = None):
        for i, layer in enumerate(self.block):
            x = layer(x)
            if taps is not None:
                taps[f"block.{i}"] = x
        return x


def vq_encode_level(q: VectorQuantize, residual: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """VectorQuantize.forward of upstream snac/vq.py (eval): returns (z_q_i at the full rate, indices, distances).
    avg_pool by the level's stride -> in_proj (768 -> 8) -> L2-normalise encodings and codebook -> nearest code
    (argmax of -dist, dist = |e|^2 - 2 e.c + |c|^2) -> un-normalised codebook vector -> out_proj -> repeat_interleave."""
    z = residual
    if q.stride > 1:
        z = F.avg_pool1d(z, q.stride, q.stride)
    z_e = q.in_proj(z)
    B, D, T = z_e.shape
    enc = z_e.permute(0, 2, 1).reshape(B * T, D)
    codebook = q.codebook.weight
    enc_n = F.normalize(enc)
    cb_n = F.normalize(codebook)
    dist = enc_n.pow(2).sum(1, keepdim=True) - 2 * enc_n @ cb_n.t() + cb_n.pow(2).sum(1, keepdim=True).t()
    indices = (-dist).max(1)[1].reshape(B, T)
    z_q = q.out_proj(q.decode_code(indices))
    if q.stride > 1:
        z_q = z_q.repeat_interleave(q.stride, dim=-1)
    return z_q, indices, dist.reshape(B, T, -1)


class SnacEncodeRef(nn.Module):
    """Encode half of ``snac.SNAC``: preprocess (right-pad to a multiple of hop * vq_strides[0] = 2048 samples),
    encoder, residual vector quantisation.  ``quantizer.quantizers`` are the same modules the decode oracle uses."""

    def __init__(self):
        super().__init__()
        self.encoder = Encoder()
        assert self.encoder.enc_dim == LATENT_DIM
        self.quantizers = nn.ModuleList(
            [VectorQuantize(LATENT_DIM, CODEBOOK_SIZE, CODEBOOK_DIM, s) for s in VQ_STRIDES])

    @staticmethod
    def preprocess(audio: torch.Tensor) -> torch.Tensor:
        length = audio.shape[-1]
        pad_to = HOP * VQ_STRIDES[0]
        right = math.ceil(length / pad_to) * pad_to - length
        return F.pad(audio, (0, right))

    @torch.inference_mode()
    def encode(self, audio: torch.Tensor, taps: Optional[dict] = None):
        """audio float [B, 1, n] -> [c0 [B, F], c1 [B, 2F], c2 [B, 4F]] int64, F = ceil(n / 2048)."""
        x = self.preprocess(audio)
        z = self.encoder(x, taps)
        if taps is not None:
            taps["z"] = z
        residual, z_q, codes = z, 0.0, []
        for i, q in enumerate(self.quantizers):
            z_q_i, idx, dist = vq_encode_level(q, residual)
            z_q = z_q + z_q_i
            residual = residual - z_q_i
            codes.append(idx)
            if taps is not None:
                taps[f"dist{i}"] = dist
                taps[f"residual{i}"] = residual
        if taps is not None:
            taps["z_q"] = z_q
        return codes

    def load_snac_state_dict(self, sd: dict):
        """Upstream key names: ``encoder.block.N...`` and ``quantizer.quantizers.N...`` (either weight-norm style)."""
        ren = {}
        for k, v in sd.items():
            k2 = k.replace("parametrizations.weight.original0", "weight_g") \
                  .replace("parametrizations.weight.original1", "weight_v")
            if k2.startswith("quantizer.quantizers."):
                k2 = k2[len("quantizer."):]
            ren[k2] = v
        own = self.state_dict()
        missing = [k for k in own if k not in ren]
        if missing:
            raise KeyError(f"checkpoint lacks encode keys: {missing[:5]} ...")
        self.load_state_dict({k: v for k, v in ren.items() if k in own}, strict=True)
        return self
