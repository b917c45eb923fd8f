"""Two-group chain kernel (kernels_chain2.cu) against the lock-step one (kernels_chain.cu) on a GPU box:
bit-identical PCM / waveform for full and sliced decodes, then per-stage times of both at B windows.

    python tests/gpu_chain2.py [B]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_inference_b200 import SnacDecoder, synth  # noqa: E402


def make(two):
    os.environ["SNACB_CHAIN2"] = "1" if two else "0"
    return SnacDecoder(synth.make_state_dict(0))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    d_old, d_new = make(False), make(True)
    ok = True
    for (b, f) in ((3, 4), (37, 4), (2, 16), (300, 4)):
        tok = torch.from_numpy(synth.make_tokens(b, f, bad_frac=0.01)).cuda()
        nz = [torch.from_numpy(n).cuda() for n in synth.make_noises(b, 4 * f, seed=3)]
        for sl in (False, True):
            if sl and f != 4:
                continue
            outs = []
            for dec in (d_old, d_new):
                pcm, wave = dec.decode(tok, raw_ids=True, noise=nz, precision="fp16", extract_slice=sl, return_wave=True)
                torch.cuda.synchronize()
                outs.append((pcm.cpu().numpy(), wave.cpu().numpy()))
            same = np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
            nbad = int((outs[0][1] != outs[1][1]).sum())
            print(f"B={b} F={f} sliced={sl}: identical={same} (differing samples {nbad}, "
                  f"max |dw| {np.abs(outs[0][1] - outs[1][1]).max():.3e}, nonzero {int((outs[1][0] != 0).sum())})", flush=True)
            ok &= same
        # in-kernel counter RNG
        outs = [dec.decode(tok, raw_ids=True, seed=11, precision="fp16").cpu().numpy() for dec in (d_old, d_new)]
        same = np.array_equal(outs[0], outs[1])
        print(f"B={b} F={f} rng: identical={same}", flush=True)
        ok &= same
    print("PARITY", "OK" if ok else "FAILED", flush=True)
    tok = torch.from_numpy(synth.make_tokens(B, 4)).cuda()
    for name, dec in (("old", d_old), ("new", d_new)):
        for sl in (False, True):
            for i in range(3):
                dec.decode(tok, raw_ids=True, extract_slice=sl, seed=i)
            dec.profile(True)
            for i in range(5):
                dec.decode(tok, raw_ids=True, extract_slice=sl, seed=i)
            rep = dec.profile_report()
            dec.profile(False)
            tot = sum(ms for _, ms in rep.values())
            print(f"[{name} sliced={sl}] total {tot / 5 * 1e3:.0f} us: " +
                  " ".join(f"{k}={ms / c * 1e3:.0f}" for k, (c, ms) in rep.items() if "chain" in k), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
