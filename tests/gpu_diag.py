"""Per-stage diagnostic on a GPU box:  python tests/gpu_diag.py [B] [F]

Prints, for the fp32 and the bf16 path, each stage's max-abs error and SNR against the oracle
(same tokens, same injected noise).  Not a test; used while bringing kernels up."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import synth_ckpt  # noqa: E402
from tests._util import oracle_decode, snr_db, pcm_of  # noqa: E402
from tts_inference_b200 import SnacDecoder, synth  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    F_ = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fp32", "bf16", "bf16s", "fp16", "fp16s"]
    sd = synth.make_state_dict(0)
    model = synth_ckpt.make_model(0, state_dict=sd)
    tokens = synth.make_tokens(B, F_, bad_frac=0.02)
    noises = synth.make_noises(B, 4 * F_, seed=7)
    t = time.time()
    ref_wave, ref_taps = oracle_decode(model, tokens, noises, want_taps=True)
    print(f"oracle: {time.time() - t:.2f}s  wave std {ref_wave.std():.4f}", flush=True)
    dec = SnacDecoder(sd)
    tok = torch.from_numpy(tokens).cuda()
    nz = [torch.from_numpy(n).cuda() for n in noises]
    for mode in modes:
        prec = mode.rstrip("s")
        try:
            pcm, wave = dec.decode(tok, raw_ids=True, noise=nz, precision=prec, return_wave=True, keep_taps=True,
                                   stream_fp32=mode.endswith("s"))
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"[{mode}] FAILED: {e}", flush=True)
            continue
        taps = dec.taps()
        print(f"[{mode}] stage                maxabs-err     SNR dB   ref-absmax")
        for k, r in ref_taps.items():
            if k not in taps:
                print(f"[{mode}] {k:18s} missing"); continue
            g = taps[k]
            if g.shape != r.shape:
                print(f"[{mode}] {k:18s} shape {g.shape} vs {r.shape}"); continue
            print(f"[{mode}] {k:18s} {np.abs(g - r).max():12.3e} {snr_db(r, g):10.2f} {np.abs(r).max():10.3f}")
        w = wave.cpu().numpy()
        p = pcm.cpu().numpy()
        print(f"[{mode}] wave               {np.abs(w - ref_wave).max():12.3e} {snr_db(ref_wave, w):10.2f}")
        print(f"[{mode}] pcm max |diff| vs oracle pcm: {np.abs(p.astype(np.int32) - pcm_of(ref_wave).astype(np.int32)).max()}"
              f"   vs own wave: {np.abs(p.astype(np.int32) - pcm_of(w).astype(np.int32)).max()}", flush=True)
    print("stats (launches, streams):", dec.stats())


if __name__ == "__main__":
    main()
