"""Throughput of the SNAC encode path (csrc/encoder.cu, fp32 CUDA cores): B utterances x F frames per call, CUDA-event
time; the oracle (PyTorch fp32, all host threads) on a bounded sample beside it.
    python tests/gpu_encode_bench.py [B] [F] > gpurun_out/encode_bench.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_inference_b200 import synth  # noqa: E402
from tts_inference_b200.encoder import SnacEncoder  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    F = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    sd = synth.make_encoder_state_dict(0)
    enc = SnacEncoder(sd)
    audio = torch.from_numpy(synth.make_audio(B, 2048 * F)).cuda()
    for _ in range(3):
        enc.encode(audio)
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); enc.encode(audio); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    audio_s = B * F * 2048 / 24000.0
    flop = 2.0 * B * F * 2048 * 100.0e3                  # ~100 kMAC per input sample (encoder FLOPs)
    from oracle.snac_enc_ref import SnacEncodeRef
    m = SnacEncodeRef().eval()
    m.load_snac_state_dict({k: torch.from_numpy(np.ascontiguousarray(v).copy()) for k, v in sd.items()})
    torch.set_num_threads(os.cpu_count() or 1)
    xs = audio[:4].cpu()[:, None, :]
    m.encode(xs)
    t0 = time.perf_counter(); m.encode(xs); dt = time.perf_counter() - t0
    print(json.dumps({
        "utterances": B, "frames": F, "ms_per_call": ms, "audio_s_per_s": audio_s / ms * 1e3, "launches_per_call": 33,
        "approx_tflops_fp32": flop / ms / 1e9,
        "cpu_oracle": {"audio_s_per_s": 4 * F * 2048 / 24000.0 / dt, "cores": os.cpu_count(), "sample": f"4 x {F} frames"},
    }, indent=1))


if __name__ == "__main__":
    main()
