"""Measurement helpers shared by bench.py and the GPU scripts (latency mode, BASELINE configs[4])."""
from __future__ import annotations

import time

import numpy as np
import torch

from . import SnacDecoder, synth


def pct(a, p):
    return float(np.percentile(np.asarray(a), p))


def measure_latency(iters=300, precision="fp16", dec=None):
    """Batch-1 single-window decode: CUDA-event time per decode (plain launches and CUDA-graph replay) and host
    wall time pinned-tokens -> pinned-int16 through snacb_decode_host."""
    if dec is None:
        dec = SnacDecoder(synth.make_state_dict(0))
    tokens = synth.make_tokens(1, 4)
    tok = torch.from_numpy(tokens).cuda()
    out = torch.empty((1, 2048), dtype=torch.int16, device="cuda")
    for i in range(10):
        dec.decode(tok, raw_ids=True, extract_slice=True, seed=i, precision=precision, out=out)
    torch.cuda.synchronize()
    res = {}
    # plain launches
    ts = []
    for i in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dec.decode(tok, raw_ids=True, extract_slice=True, seed=i, precision=precision, out=out)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    res["device_ms_plain"] = {"p50": pct(ts, 50), "p99": pct(ts, 99)}
    # CUDA graph replay
    try:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            dec.decode(tok, raw_ids=True, extract_slice=True, seed=1, precision=precision, out=out)
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(g):
            dec.decode(tok, raw_ids=True, extract_slice=True, seed=1, precision=precision, out=out)
        ref = out.clone()
        ts = []
        for i in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        assert torch.equal(ref, out)
        res["device_ms_graph"] = {"p50": pct(ts, 50), "p99": pct(ts, 99)}
    except Exception as e:  # noqa: BLE001
        res["device_ms_graph"] = {"error": str(e)[:200]}
    # host-to-host wall time
    tp = torch.from_numpy(tokens).pin_memory()
    pp = torch.empty((1, 2048), dtype=torch.int16).pin_memory()
    for i in range(10):
        dec.decode_host_ptr(tp.data_ptr(), 1, 28, pp.data_ptr(), raw_ids=True, extract_slice=True, seed=i, precision=precision)
    ts = []
    for i in range(iters):
        t = time.perf_counter()
        dec.decode_host_ptr(tp.data_ptr(), 1, 28, pp.data_ptr(), raw_ids=True, extract_slice=True, seed=i, precision=precision)
        ts.append((time.perf_counter() - t) * 1e3)
    res["host_ms_e2e"] = {"p50": pct(ts, 50), "p99": pct(ts, 99)}
    res["precision"] = precision
    l0 = dec.stats()[0]
    dec.decode(tok, raw_ids=True, extract_slice=True, seed=0, precision=precision, out=out)
    res["launches_per_decode"] = dec.stats()[0] - l0          # counted, not assumed
    return res


