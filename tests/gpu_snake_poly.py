"""Half2-polynomial Snake in the fp16 chain prologue (SNACB_SNAKE_POLY = 0 / 2, kernels_chain.cu::span_half; the build
measured in DESIGN.md section 6 also instantiated 1 = snake1 only and 3 = both)
on a GPU box: SNR of each setting against the oracle (same tokens, same injected noise), then the chain kernels'
times at B windows.

    python tests/gpu_snake_poly.py [B]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth_ckpt  # noqa: E402  (checker only)
from tests._util import oracle_decode, snr_db  # noqa: E402
from tts_inference_b200 import SnacDecoder, synth  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    sd = synth.make_state_dict(0)
    decs = {}
    for p in (0, 2):
        os.environ["SNACB_SNAKE_POLY"] = str(p)
        decs[p] = SnacDecoder(sd)
    os.environ.pop("SNACB_SNAKE_POLY")
    model = synth_ckpt.make_model(0, state_dict=sd)
    for (b, f) in ((4, 4), (2, 16)):
        tokens = synth.make_tokens(b, f, bad_frac=0.01)
        noises = synth.make_noises(b, 4 * f, seed=5)
        ref, _ = oracle_decode(model, tokens, noises)
        tok = torch.from_numpy(tokens).cuda()
        nz = [torch.from_numpy(n).cuda() for n in noises]
        for p, dec in decs.items():
            _, w = dec.decode(tok, raw_ids=True, noise=nz, precision="fp16", return_wave=True)
            torch.cuda.synchronize()
            print(f"B={b} F={f} poly={p}: SNR {snr_db(ref, w.cpu().numpy()):.2f} dB", flush=True)
    tok = torch.from_numpy(synth.make_tokens(B, 4)).cuda()
    for p, dec in decs.items():
        for sl in (False, True):
            for i in range(3):
                dec.decode(tok, raw_ids=True, extract_slice=sl, seed=i)
            dec.profile(True)
            for i in range(5):
                dec.decode(tok, raw_ids=True, extract_slice=sl, seed=i)
            rep = dec.profile_report()
            dec.profile(False)
            tot = sum(ms for _, ms in rep.values())
            print(f"[poly={p} sliced={sl}] total {tot / 5 * 1e3:.0f} us: " +
                  " ".join(f"{k}={ms / c * 1e3:.0f}" for k, (c, ms) in rep.items() if "chain" in k), flush=True)


if __name__ == "__main__":
    main()
