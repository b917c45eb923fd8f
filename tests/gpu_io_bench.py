"""Timing of the ingest / egress kernels on a GPU box (CUDA events, L2 flushed between repetitions):

    python tests/gpu_io_bench.py > gpurun_out/io_bench.json
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_inference_b200 import egress  # noqa: E402
from tts_inference_b200.ingest import DeviceIngest  # noqa: E402


def timed(fn, reps=20, flush=None):
    ms = []
    for _ in range(3):
        fn()
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def main():
    peaks = {}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {"peaks_file": peaks}
    n, samples = 1024, 8192
    pcm = torch.randint(-32768, 32767, (n, samples), dtype=torch.int16, device="cuda")
    b64 = torch.empty((n, 4 * ((2 * samples + 2) // 3)), dtype=torch.uint8, device="cuda")
    wav = torch.empty((n, 44 + 2 * samples), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: egress.pcm_to_base64(pcm, out=b64), flush=flush)
    by = pcm.numel() * 2 + b64.numel()
    out["base64"] = {"chunks": n, "samples": samples, "ms": ms, "bytes": by, "GB_per_s": by / ms / 1e6}
    ms = timed(lambda: egress.pcm_to_wav(pcm, out=wav), flush=flush)
    by = pcm.numel() * 2 + wav.numel()
    out["wav"] = {"records": n, "samples": samples, "ms": ms, "bytes": by, "GB_per_s": by / ms / 1e6}
    for S, k in ((1024, 1), (4096, 1), (4096, 7), (4096, 28)):
        ing = DeviceIngest(S)
        base = 128266
        tok = (base + torch.randint(0, 28672, (S, k), dtype=torch.int32, device="cuda")).contiguous()
        sos = torch.full((S, k), 128257, dtype=torch.int32, device="cuda")
        ing.step(sos)
        lib, g = ing._lib, ing._g
        cap = int(lib.snacb_ingest_window_capacity(S, k))
        wt = torch.empty((cap, 28), dtype=torch.int32, device="cuda")
        ws = torch.empty(cap, dtype=torch.int32, device="cuda")

        def one():
            lib.snacb_ingest_step(g, tok.data_ptr(), S, k, None, None, wt.data_ptr(), ws.data_ptr(), cap,
                                  ing._tail_tok.data_ptr(), ing._tail_stream.data_ptr(), ing._tail_frames.data_ptr(),
                                  ing._counts.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ms = timed(one, reps=56)
        out[f"ingest_S{S}_n{k}"] = {"us_per_step": ms * 1e3, "tokens_per_s": S * k / ms * 1e3}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
