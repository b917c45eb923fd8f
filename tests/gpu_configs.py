"""All five BASELINE.json configs on one GPU box (not a test; writes gpurun_out/configs.json):

  1. one 28-token window, batch 1, fp32                    (parity config; time per decode)
  2. 256 / 1024 concurrent streams x 28-token windows, fp16 tensor-core path
  3. full-utterance decode, batch 64, F in {16, 64, 512} frames (512 frames = 43.7 s of audio per stream)
  4. "Hindi vocabulary" streams: 0.5 % out-of-range / wrong-position ids, 512 streams per GPU (of 4096 over 8)
  5. latency mode: batch 1 single window, CUDA graph, p50 / p99
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_inference_b200 import SnacDecoder, synth  # noqa: E402
from tts_inference_b200.bench_util import measure_latency  # noqa: E402


def timed(fn, reps):
    fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i + 1)
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dec = SnacDecoder(synth.make_state_dict(0))
    out = {}
    tok1 = torch.from_numpy(synth.make_tokens(1, 4)).cuda()
    ms = timed(lambda i: dec.decode(tok1, raw_ids=True, seed=i, precision="fp32"), 20)
    out["cfg1_b1_fp32_window"] = {"ms_per_decode": ms, "audio_s_per_s": 8192 / 24000 / (ms * 1e-3)}
    for B in (256, 1024):
        tok = torch.from_numpy(synth.make_tokens(B, 4, seed=20241224)).cuda()
        full = timed(lambda i: dec.decode(tok, raw_ids=True, seed=i), 10)
        sl = timed(lambda i: dec.decode(tok, raw_ids=True, seed=i, extract_slice=True), 10)
        out[f"cfg2_b{B}_windows_fp16"] = {
            "full_window_ms": full, "decoded_audio_s_per_s": B * 8192 / 24000 / (full * 1e-3),
            "sliced_ms": sl, "sliced_windows_per_s": B / (sl * 1e-3), "sliced_emitted_audio_s_per_s": B * 2048 / 24000 / (sl * 1e-3)}
    for F_ in (16, 64, 512):
        B = 64
        tok = torch.from_numpy(synth.make_tokens(B, F_, seed=7)).cuda()
        reps = 3 if F_ == 512 else 5
        ms = timed(lambda i: dec.decode(tok, raw_ids=True, seed=i), reps)
        out[f"cfg3_b64_F{F_}_utterance_fp16"] = {"ms": ms, "audio_s_per_stream": F_ * 2048 / 24000,
                                                  "decoded_audio_s_per_s": B * F_ * 2048 / 24000 / (ms * 1e-3)}
    tokh = synth.make_tokens(512, 4, seed=99, bad_frac=0.005)
    tok = torch.from_numpy(tokh).cuda()
    c = dec.unpack(tok, raw_ids=True)
    assert all(int(x.min()) >= 0 and int(x.max()) <= 4095 for x in c)
    ms = timed(lambda i: dec.decode(tok, raw_ids=True, seed=i, extract_slice=True), 10)
    out["cfg4_hindi_vocab_512_streams_per_gpu"] = {"sliced_ms": ms, "windows_per_s": 512 / (ms * 1e-3),
                                                   "bad_id_fraction": 0.005, "codes_in_range": True}
    out["cfg5_latency_b1"] = measure_latency(300, "fp16", dec)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/configs.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
