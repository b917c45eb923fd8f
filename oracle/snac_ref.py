"""PyTorch fp32 restatement of ``snac.SNAC.decode`` for the 24 kHz checkpoint (ORACLE).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  PARITY UNPINNED at this boundary:
the upstream package ``snac`` (PyPI, github hubertsiuzdak/snac; the reference installs it
unpinned, vllm_inference/modal_audio_stream.py:58) is not in this image, so this file
restates its published module graph.  Reference call sites that enter this code:
``snac_model.decode(codes)`` vllm_inference/modal_audio_stream.py:191,
tensorrt_tts/inference.py:106, tensorrt_tts/hindi_canopy/inference.py:195,
tensorrt_tts/hindi_finetuned/benchmark.py:232.

Module tree and state-dict keys are the upstream ones (``quantizer.quantizers.N.*``,
``decoder.model.N.*``; weight-norm in the old ``weight_g`` / ``weight_v`` form), so a real
``pytorch_model.bin`` loads with ``load_state_dict`` when someone supplies one
(new-style ``parametrizations.weight.original0/1`` keys are renamed on load).

The one deliberate difference: ``NoiseBlock`` accepts an injected noise tensor so that
the CUDA path and the oracle can be fed identical noise (north_star's parity rule).
"""
from __future__ import annotations

import math
import warnings
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

# snac_24khz/config.json
SAMPLING_RATE = 24000
LATENT_DIM = 768            # encoder_dim 48 * 2**4
DECODER_DIM = 1024
DECODER_RATES = (8, 8, 4, 2)
CODEBOOK_SIZE = 4096
CODEBOOK_DIM = 8
VQ_STRIDES = (4, 2, 1)
HOP = 512                   # samples per latent step


def _wn(module: nn.Module) -> nn.Module:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.nn.utils.weight_norm(module)


def WNConv1d(*a, **k):
    return _wn(nn.Conv1d(*a, **k))


def WNConvTranspose1d(*a, **k):
    return _wn(nn.ConvTranspose1d(*a, **k))


def snake(x: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """x + (alpha + 1e-9)^-1 * sin(alpha x)^2   (upstream snac/layers.py ``snake``)."""
    shape = x.shape
    x = x.reshape(shape[0], shape[1], -1)
    x = x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)
    return x.reshape(shape)


class Snake1d(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.alpha = nn.Parameter(torch.ones(1, channels, 1))

    def forward(self, x):
        return snake(x, self.alpha)


class NoiseBlock(nn.Module):
    """x + randn(B,1,T) * Conv1x1_nobias(x)   (upstream snac/layers.py ``NoiseBlock``)."""

    def __init__(self, dim: int):
        super().__init__()
        self.linear = WNConv1d(dim, dim, kernel_size=1, bias=False)

    def forward(self, x, noise: Optional[torch.Tensor] = None):
        B, C, T = x.shape
        if noise is None:
            noise = torch.randn((B, 1, T), device=x.device, dtype=x.dtype)
        assert noise.shape == (B, 1, T), (noise.shape, (B, 1, T))
        h = self.linear(x)
        return x + noise * h


class ResidualUnit(nn.Module):
    def __init__(self, dim: int, dilation: int, kernel: int = 7, groups: int = 1):
        super().__init__()
        pad = ((kernel - 1) * dilation) // 2
        self.block = nn.Sequential(
            Snake1d(dim),
            WNConv1d(dim, dim, kernel_size=kernel, dilation=dilation, padding=pad, groups=groups),
            Snake1d(dim),
            WNConv1d(dim, dim, kernel_size=1),
        )

    def forward(self, x):
        y = self.block(x)
        pad = (x.shape[-1] - y.shape[-1]) // 2
        if pad > 0:
            x = x[..., pad:-pad]
        return x + y


class DecoderBlock(nn.Module):
    def __init__(self, input_dim: int, output_dim: int, stride: int, noise: bool, groups: int):
        super().__init__()
        layers: List[nn.Module] = [
            Snake1d(input_dim),
            WNConvTranspose1d(
                input_dim, output_dim, kernel_size=2 * stride, stride=stride,
                padding=math.ceil(stride / 2), output_padding=stride % 2,
            ),
        ]
        if noise:
            layers.append(NoiseBlock(output_dim))
        layers += [
            ResidualUnit(output_dim, dilation=1, groups=groups),
            ResidualUnit(output_dim, dilation=3, groups=groups),
            ResidualUnit(output_dim, dilation=9, groups=groups),
        ]
        self.block = nn.Sequential(*layers)

    def forward(self, x, noise: Optional[torch.Tensor] = None, taps: Optional[dict] = None, name: str = ""):
        for i, layer in enumerate(self.block):
            if isinstance(layer, NoiseBlock):
                x = layer(x, noise)
            else:
                x = layer(x)
            if taps is not None:
                taps[f"{name}.{i}"] = x
        return x


class Decoder(nn.Module):
    def __init__(self, input_channel=LATENT_DIM, channels=DECODER_DIM, rates=DECODER_RATES,
                 noise=True, depthwise=True, d_out=1):
        super().__init__()
        assert depthwise, "snac_24khz uses depthwise=True"
        layers: List[nn.Module] = [
            WNConv1d(input_channel, input_channel, kernel_size=7, padding=3, groups=input_channel),
            WNConv1d(input_channel, channels, kernel_size=1),
        ]
        output_dim = channels
        for i, stride in enumerate(rates):
            input_dim = channels // 2 ** i
            output_dim = channels // 2 ** (i + 1)
            layers.append(DecoderBlock(input_dim, output_dim, stride, noise, groups=output_dim))
        layers += [Snake1d(output_dim), WNConv1d(output_dim, d_out, kernel_size=7, padding=3), nn.Tanh()]
        self.model = nn.Sequential(*layers)

    def forward(self, z, noises: Optional[Sequence[torch.Tensor]] = None, taps: Optional[dict] = None):
        x = z
        bi = 0
        for i, layer in enumerate(self.model):
            if isinstance(layer, DecoderBlock):
                n = None if noises is None else noises[bi]
                x = layer(x, n, taps, f"model.{i}")
                bi += 1
            else:
                x = layer(x)
            if taps is not None:
                taps[f"model.{i}"] = x
        return x


class VectorQuantize(nn.Module):
    def __init__(self, input_dim: int, codebook_size: int, codebook_dim: int, stride: int):
        super().__init__()
        self.stride = stride
        self.in_proj = WNConv1d(input_dim, codebook_dim, kernel_size=1)   # unused by decode
        self.out_proj = WNConv1d(codebook_dim, input_dim, kernel_size=1)
        self.codebook = nn.Embedding(codebook_size, codebook_dim)

    def decode_code(self, embed_id):
        return F.embedding(embed_id, self.codebook.weight).transpose(1, 2)


class ResidualVectorQuantize(nn.Module):
    def __init__(self, input_dim=LATENT_DIM, codebook_size=CODEBOOK_SIZE, codebook_dim=CODEBOOK_DIM,
                 vq_strides=VQ_STRIDES):
        super().__init__()
        self.n_codebooks = len(vq_strides)
        self.quantizers = nn.ModuleList(
            [VectorQuantize(input_dim, codebook_size, codebook_dim, s) for s in vq_strides])

    def from_codes(self, codes: Sequence[torch.Tensor]) -> torch.Tensor:
        z_q = 0.0
        for i in range(self.n_codebooks):
            z_p_i = self.quantizers[i].decode_code(codes[i])
            z_q_i = self.quantizers[i].out_proj(z_p_i)
            z_q_i = z_q_i.repeat_interleave(self.quantizers[i].stride, dim=-1)
            z_q = z_q + z_q_i
        return z_q


class SnacDecodeRef(nn.Module):
    """Decode half of ``snac.SNAC`` (quantizer.from_codes + decoder)."""

    def __init__(self):
        super().__init__()
        self.quantizer = ResidualVectorQuantize()
        self.decoder = Decoder()

    def noise_shapes(self, batch: int, t0: int):
        """Shapes of the four injected NoiseBlock tensors for ``t0`` latent steps."""
        out, t = [], t0
        for r in DECODER_RATES:
            t *= r
            out.append((batch, 1, t))
        return out

    @torch.inference_mode()
    def decode(self, codes: Sequence[torch.Tensor], noises: Optional[Sequence[torch.Tensor]] = None,
               taps: Optional[dict] = None) -> torch.Tensor:
        z_q = self.quantizer.from_codes(codes)
        if taps is not None:
            taps["z_q"] = z_q
        return self.decoder(z_q, noises, taps)

    def load_snac_state_dict(self, sd: dict, strict_decode: bool = True):
        """Load an upstream ``snac`` state dict (old or new weight-norm key style);
        encoder / in_proj keys are ignored."""
        ren = {}
        for k, v in sd.items():
            k2 = k.replace("parametrizations.weight.original0", "weight_g") \
                  .replace("parametrizations.weight.original1", "weight_v")
            ren[k2] = v
        own = self.state_dict()
        missing = [k for k in own if k not in ren and ".in_proj." not in k]
        if strict_decode and missing:
            raise KeyError(f"checkpoint lacks decode keys: {missing[:5]} ...")
        self.load_state_dict({k: v for k, v in ren.items() if k in own}, strict=False)
        return self
