"""CPU emulation of where the tensor-core path rounds to 16 bits (storage between kernels, MMA operands, the chain
kernel's tile copy), to predict the SNR of a precision scheme before building it.  Test infrastructure (imports the
oracle).   python tools/precision_emul.py"""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import glue_ref, synth_ckpt
from oracle.snac_ref import snake
from tts_inference_b200 import synth


def rnd(x, dt):
    if dt is None:
        return x
    return x.to(dt).to(torch.float32)


def split(w, dt, parts):
    """w as a sum of `parts` 16-bit values (what parts MMAs against the same A compute)."""
    acc = torch.zeros_like(w)
    r = w.clone()
    for _ in range(parts):
        p = rnd(r, dt)
        acc = acc + p
        r = r - p
    return acc


def wfold(conv):
    return torch._weight_norm(conv.weight_v, conv.weight_g, 0)


@torch.inference_mode()
def decode_emul(m, codes, noises, dt, chain_blocks=(1, 2, 3), w_parts_full=1, w_parts_res=1, store_dt=None, a_parts=1,
                fp32_stream_unfused=False):
    """dt: operand dtype; store_dt: dtype of tensors stored between kernels (default dt); w_parts_full: weight split of the
    stem / ConvTranspose GEMMs; w_parts_res: of the NoiseBlock / ResidualUnit 1x1s; a_parts: split of the chain's MMA operand."""
    sdt = store_dt or dt
    z = m.quantizer.from_codes(codes)
    dec = m.decoder.model
    a0 = rnd(dec[0](z), sdt)
    w = split(wfold(dec[1]), dt, w_parts_full)
    x = F.conv1d(rnd(a0, dt), w, dec[1].bias)
    for bi in range(4):
        blk = dec[2 + bi].block
        x = rnd(snake(x, blk[0].alpha), sdt)                         # producer epilogue: Snake + store
        ct = blk[1]
        w = split(wfold(ct), dt, w_parts_full)
        y = F.conv_transpose1d(rnd(x, dt), w, ct.bias, stride=ct.stride, padding=ct.padding, output_padding=ct.output_padding)
        y = rnd(y, sdt)
        fused = bi in chain_blocks
        wn = split(wfold(blk[2].linear), dt, w_parts_res)
        x = y + noises[bi] * F.conv1d(rnd(y, dt), wn)
        if not fused and not fp32_stream_unfused:
            x = rnd(x, sdt)
        for ri in range(3):
            ru = blk[3 + ri].block
            s1 = snake(x, ru[0].alpha)
            if fused:
                s1 = rnd(s1, sdt)                                    # the tile copy
            h = F.conv1d(s1, wfold(ru[1]), ru[1].bias, dilation=ru[1].dilation, padding=ru[1].padding, groups=ru[1].groups)
            a = snake(h, ru[2].alpha)
            a = split(a, dt, a_parts)
            wp = split(wfold(ru[3]), dt, w_parts_res)
            x = x + F.conv1d(a, wp, ru[3].bias)
            if not fused and not fp32_stream_unfused and ri < 2:
                x = rnd(x, sdt)
    x = rnd(snake(x, dec[6].alpha), sdt)
    return torch.tanh(dec[7](x))


def snr(ref, got):
    return float(10 * torch.log10((ref ** 2).sum() / ((ref - got) ** 2).sum()))


if __name__ == "__main__":
    m = synth_ckpt.make_model(0)
    tokens = synth.make_tokens(4, 4, seed=11)
    lv = glue_ref.unpack_np(tokens.astype(np.int64) - 128266)
    codes = [torch.from_numpy(x.astype(np.int64)) for x in lv]
    noises = [torch.from_numpy(n) for n in synth.make_noises(4, 16, seed=7)]
    with torch.inference_mode():
        ref = m.decode(codes, noises)
    bf, hf = torch.bfloat16, torch.float16
    print("fp32 emul            ", snr(ref, decode_emul(m, codes, noises, None)))
    print("fp16 (today)         ", snr(ref, decode_emul(m, codes, noises, hf)))
    print("bf16 (today, b1 unf.)", snr(ref, decode_emul(m, codes, noises, bf, chain_blocks=(2, 3))))
    print("bf16 all chains      ", snr(ref, decode_emul(m, codes, noises, bf)))
    print("bf16 W2 full GEMMs   ", snr(ref, decode_emul(m, codes, noises, bf, w_parts_full=2)))
    print("bf16 W2 everywhere   ", snr(ref, decode_emul(m, codes, noises, bf, w_parts_full=2, w_parts_res=2)))
    print("bf16 W2 + A2 chain   ", snr(ref, decode_emul(m, codes, noises, bf, w_parts_full=2, w_parts_res=2, a_parts=2)))
    print("bf16 ops, fp16 store ", snr(ref, decode_emul(m, codes, noises, bf, store_dt=hf)))
    print("bf16 ops W2, fp16 st ", snr(ref, decode_emul(m, codes, noises, bf, store_dt=hf, w_parts_full=2, w_parts_res=2)))
    print("bf16 W2+A2, fp16 st  ", snr(ref, decode_emul(m, codes, noises, bf, store_dt=hf, w_parts_full=2, w_parts_res=2, a_parts=2)))
    print("bf16 W2, b1 unfused, fp32 stream", snr(ref, decode_emul(m, codes, noises, bf, chain_blocks=(2, 3), w_parts_full=2, w_parts_res=2, fp32_stream_unfused=True)))


@torch.inference_mode()
def decode_emul_unfused(m, codes, noises, op_dt, st_dt, w_parts, a_parts_gemm, a_parts_res):
    """Every layer its own kernel: outputs stored in st_dt; GEMM A operands (loaded from storage) and ResidualUnit A operands
    (built in registers) split into a_parts_* values of op_dt, weights into w_parts."""
    z = m.quantizer.from_codes(codes)
    dec = m.decoder.model
    a0 = rnd(dec[0](z), st_dt)
    x = F.conv1d(split(a0, op_dt, a_parts_gemm), split(wfold(dec[1]), op_dt, w_parts), dec[1].bias)
    for bi in range(4):
        blk = dec[2 + bi].block
        x = rnd(snake(x, blk[0].alpha), st_dt)
        ct = blk[1]
        y = F.conv_transpose1d(split(x, op_dt, a_parts_gemm), split(wfold(ct), op_dt, w_parts), ct.bias, stride=ct.stride,
                               padding=ct.padding, output_padding=ct.output_padding)
        y = rnd(y, st_dt)
        x = rnd(y + noises[bi] * F.conv1d(split(y, op_dt, a_parts_gemm), split(wfold(blk[2].linear), op_dt, w_parts)), st_dt)
        for ri in range(3):
            ru = blk[3 + ri].block
            h = F.conv1d(snake(x, ru[0].alpha), wfold(ru[1]), ru[1].bias, dilation=ru[1].dilation, padding=ru[1].padding,
                         groups=ru[1].groups)
            a = split(snake(h, ru[2].alpha), op_dt, a_parts_res)
            x = x + F.conv1d(a, split(wfold(ru[3]), op_dt, w_parts), ru[3].bias)
            if ri < 2:
                x = rnd(x, st_dt)
    x = rnd(snake(x, dec[6].alpha), st_dt)
    return torch.tanh(dec[7](x))


if __name__ == "__main__":
    # the bf16x3 path that was built is the (2, 2, 1) row: predicted 47.72 dB, measured 47.74 dB on B200
    print("--- unfused, fp16 storage, bf16 MMA operands")
    for wp, ag, ar in ((1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2), (1, 2, 2), (2, 1, 2)):
        print(f"W x{wp}, GEMM A x{ag}, res A x{ar}:", snr(ref, decode_emul_unfused(m, codes, noises, bf, hf, wp, ag, ar)))
