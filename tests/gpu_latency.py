"""Latency mode (BASELINE configs[4]): batch-1 single-window decode, p50/p99 per chunk.

  python tests/gpu_latency.py [iters]

Reports (a) kernel-only CUDA-event time per decode, plain launches and CUDA-graph replay, and
(b) host wall time from tokens in pinned host memory to int16 in pinned host memory
(snacb_decode_host, the boundary the reference's helper has)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_inference_b200.bench_util import measure_latency as measure  # noqa: E402


if __name__ == "__main__":
    it = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    print(json.dumps({"latency_b1_window": measure(it, "fp16"), "latency_b1_window_fp32": measure(max(it // 3, 20), "fp32")}))
