"""Throughput of the window batcher fed by several producer threads:

    python tests/gpu_batcher_bench.py [streams] [threads] > gpurun_out/batcher_bench.json

Producer threads (native: tests/native/batcher_bench.cpp, the GIL is not involved) push the token streams of `streams`
concurrent requests 1 / 7 tokens at a time while the main thread flushes every ready window through the pipelined flush
(snacb_batcher_flush_submit / _wait) into pinned host buffers.  Reported beside the raw pipelined host call
(snacb_decode_host_submit / _wait) on the same batch size: the batcher should not cost the GPU any throughput."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tts_inference_b200 import SnacDecoder, synth  # noqa: E402
from tts_inference_b200.batcher import POLICY_CHUNK, POLICY_SLIDING, WindowBatcher  # noqa: E402


def raw_rate(dec, B, sliced, steps=12):
    tok = torch.from_numpy(synth.make_tokens(B, 4)).pin_memory()
    ns = 2048 if sliced else 8192
    outs = [torch.empty((B, ns), dtype=torch.int16).pin_memory() for _ in range(2)]
    for i in range(2):
        dec.submit_host_ptr(tok.data_ptr(), B, 28, outs[i].data_ptr(), raw_ids=True, extract_slice=sliced, seed=i)
    dec.wait_host(); dec.wait_host()
    t0 = time.perf_counter()
    for i in range(steps):
        dec.submit_host_ptr(tok.data_ptr(), B, 28, outs[i & 1].data_ptr(), raw_ids=True, extract_slice=sliced, seed=10 + i)
        if i:
            dec.wait_host()
    dec.wait_host()
    return B * steps / (time.perf_counter() - t0)


def main():
    streams = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    lib = C.CDLL(os.path.join(ROOT, "tests", "native", "libbatcherbench.so"))
    lib.batcher_bench.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_size_t, C.POINTER(C.c_double)]
    dec = SnacDecoder(synth.make_state_dict(0))
    res = {"streams": streams, "producer_threads": threads, "host_cores": os.cpu_count(), "runs": []}
    for policy, name, sliced in ((POLICY_CHUNK, "chunk (28 codes -> 8192 samples)", False),
                                 (POLICY_SLIDING, "sliding (every 7 codes -> samples [2048:4096])", True)):
        B = streams
        raw = raw_rate(dec, B, sliced)
        for push in (1, 7):
            b = WindowBatcher(dec, policy=policy, raw_ids=True, max_windows=B)
            per = 2048 if sliced else 8192
            pins = [torch.empty(B * per, dtype=torch.int16).pin_memory() for _ in range(2)]
            tokens = 28 * (12 if not sliced else 4)            # per stream: 12 chunks, or 28 + 84 tokens -> 13 windows
            if sliced:
                tokens = 28 + 7 * 12
            out = (C.c_double * 4)()
            for rep in range(2):                               # first repetition warms the workspace up
                b2 = b if rep == 0 else WindowBatcher(dec, policy=policy, raw_ids=True, max_windows=B)
                rc = lib.batcher_bench(b2._b, streams, tokens, threads, push, B, pins[0].data_ptr(), pins[1].data_ptr(),
                                       pins[0].numel(), out)
                assert rc == 0, rc
                if rep:
                    b2.close()
            b.close()
            secs, windows, pushes, flushes = out[0], out[1], out[2], out[3]
            res["runs"].append({"policy": name, "tokens_per_push": push, "windows": int(windows), "pushes": int(pushes),
                                "flushes": int(flushes), "seconds": secs, "windows_per_s": windows / secs,
                                "pushes_per_s": pushes / secs, "raw_pipelined_decode_windows_per_s": raw,
                                "batcher_over_raw": windows / secs / raw})
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
