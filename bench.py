#!/usr/bin/env python
"""Benchmark of the SNAC-24 kHz decode hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the hot path over one batch of synthetic token windows: B concurrent
streams x one 28-token (4-frame) window each -> unpack -> VQ -> decoder -> all 8192 int16 samples of
the window (convert_to_audio with extract_slice=False, the call the reference's stream_audio makes).
The sliding-window call (extract_slice=True: only samples [2048:4096] are emitted, and only their
receptive field is computed) is measured beside it and reported under "sliding_window_mode".  N>1 runs under torchrun, one rank per GPU, streams sharded with NO data-path
collective (weak scaling: B windows per GPU); NCCL is used only for the barrier and the
max-over-ranks of the device time.  Rank 0 prints ONE JSON line.

metric  : audio-sec decoded/sec = windows/s x 8192/24000; every sample of every window is computed
          and written (no dead-sample trimming in the headline)
value   : tokens already in HBM, CUDA-event time of K steps (L2 flushed between steps)
e2e     : same metric through the host-buffer C-ABI call (pinned host tokens -> H2D -> decode ->
          D2H -> sync every step) -- the boundary the reference's convert_to_audio has
roofline: dominant kernel class by share of step time, algorithmic FLOPs / CUDA-event duration vs
          the measured sustained bf16 tensor peak (MEASURED_PEAKS.json)
cpu_baseline / --impl reference: the oracle port of the reference's PyTorch SNAC decoder on the
          host cores (the pip package `snac` is not installable here).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "audio-sec decoded/sec (SNAC 24kHz)"
UNIT = "audio-s/s"
FRAMES = 4
WINDOW_SAMPLES = 2048 * FRAMES
SR = 24000.0
FLOP_PER_LATENT_STEP = 207.0e6          # 103.50 M MAC per latent step
FLOP_PER_WINDOW = 16 * FLOP_PER_LATENT_STEP


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1389.5), "hbm_gbs": d.get("hbm_gbs", 6552.6),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------------
# per-stage algorithmic FLOPs for ONE group launch of S windows (F=4), used for the roofline
# ----------------------------------------------------------------------------------------------
def stage_flops(S: int) -> dict:
    out = {}
    t0 = 16
    out["stem_pw"] = 2.0 * S * t0 * 768 * 1024
    cin, t = 1024, t0
    for bi, s in enumerate((8, 8, 4, 2)):
        cout = cin // 2
        tout = t * s
        out[f"b{bi}.convt"] = 2.0 * S * tout * 2 * cin * cout          # 2 taps per output sample
        out[f"b{bi}.noise"] = 2.0 * S * tout * cout * cout
        for ri in range(3):
            out[f"b{bi}.res{ri}"] = 2.0 * S * tout * (cout * cout + 7 * cout)
        # fused NoiseBlock + 3 ResidualUnits (k_chain): same algorithmic work in one launch
        out[f"b{bi}.chain"] = out[f"b{bi}.noise"] + 3 * out[f"b{bi}.res0"]
        cin, t = cout, tout
    out["tail"] = 2.0 * S * 8192 * 64 * 7
    out["vq_stem"] = 2.0 * S * t0 * 768 * (24 + 7)
    out["unpack"] = 0.0
    return out


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (pynvml)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake": 0x80}
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=1)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_port_rate(batch: int, reps: int, threads: int, state_dict=None):
    """Oracle port of the reference decoder on the host cores: decoded audio-s/s, protocol of
    tensorrt_tts/hindi_finetuned/benchmark.py:219-247 (tensor construction + clamp + decode + int16)."""
    import torch
    from oracle import glue_ref, synth_ckpt
    from tts_inference_b200 import synth
    torch.set_num_threads(threads)
    model = synth_ckpt.make_model(0, state_dict=state_dict)
    tokens = synth.make_tokens(batch, FRAMES)
    codes = tokens.astype(np.int64) - 128266

    def one():
        lv = glue_ref.unpack_np(codes)
        y = model.decode([torch.from_numpy(x.astype(np.int64)) for x in lv])
        return glue_ref.pcm16_torch(y)

    one()
    t = time.perf_counter()
    for _ in range(reps):
        one()
    dt = (time.perf_counter() - t) / reps
    return batch * WINDOW_SAMPLES / SR / dt, dt


def gpu_torch_baseline():
    """The reference-equivalent PyTorch module (oracle restatement of snac.SNAC.decode: eager torch, cuDNN convs,
    weight-norm recomputed per forward, TF32 defaults, cudnn.benchmark=True as init_snac sets it,
    modal_audio_stream.py:116) on THIS GPU, protocol of tensorrt_tts/hindi_finetuned/benchmark.py:219-247: the timer
    spans torch.tensor(list) -> clamp -> decode -> synchronize.  The closest stand-in for the reference's own GPU
    numbers (the pip package / checkpoint cannot be installed here).  Reported beside the headline, never inside it."""
    import torch
    from oracle import glue_ref, synth_ckpt
    from tts_inference_b200 import synth
    torch.backends.cudnn.benchmark = True
    model = synth_ckpt.make_model(0).cuda().eval()
    out = {"what": "oracle restatement of the reference's PyTorch SNAC decoder, eager fp32/TF32 on this GPU, cudnn.benchmark, "
                   "timer = tensor build + clamp + decode + sync (benchmark.py:219-247); 5 warm-up + 20 timed, median"}

    def run(batch, frames):
        codes = synth.make_tokens(batch, frames).astype(np.int64) - 128266
        lv = [x.tolist() for x in glue_ref.unpack_np(codes)]

        def once():
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ct = [torch.tensor(l, dtype=torch.int32, device="cuda") for l in lv]
            ct = [torch.clamp(c, 0, 4095) for c in ct]
            with torch.inference_mode():
                model.decode(ct)
            torch.cuda.synchronize()
            return time.perf_counter() - t0
        for _ in range(5):
            once()
        ts = sorted(once() for _ in range(20))
        med = ts[len(ts) // 2]
        return {"ms": med * 1e3, "audio_s_per_s": batch * frames * 2048 / SR / med}
    out["b1_window"] = run(1, 4)            # what convert_to_audio does per call: one stream, one 28-token window
    out["b64_windows"] = run(64, 4)         # batching the reference never does, for scale
    out["b1_utterance_512_frames"] = run(1, 512)
    del model
    torch.cuda.empty_cache()
    return out


def extra_configs(dec, args, world, max_over_ranks):
    """BASELINE.json configs 1, 3 and 4 (config 2 is the headline, config 5 the latency block).  CUDA-event times, max over
    ranks; every rank works on its own streams (no collective)."""
    import torch
    from tts_inference_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    res = {}

    def timed(fn, reps, warm=2):
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i in range(reps):
            ev[i][0].record(); fn(100 + i); ev[i][1].record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        return max_over_ranks(ts[len(ts) // 2], device="cuda")
    # config 1: one 28-token window, batch 1, fp32 CUDA-core path
    t1 = torch.from_numpy(synth.make_tokens(1, FRAMES, seed=20241224 + rank)).cuda()
    o1 = torch.empty((1, WINDOW_SAMPLES), dtype=torch.int16, device="cuda")
    ms = timed(lambda i: dec.decode(t1, raw_ids=True, seed=i, precision="fp32", out=o1), 30, 3)
    res["config1_b1_window_fp32"] = {"ms": ms, "audio_s_per_s": WINDOW_SAMPLES / SR / (ms * 1e-3)}
    # config 3: full utterances, 64 streams x F frames (512 frames = 43.7 s of audio per stream)
    sweep = {}
    for F in (16, 64, 512):
        tk = torch.from_numpy(synth.make_tokens(64, F, seed=20241224 + rank)).cuda()
        ok = torch.empty((64, 2048 * F), dtype=torch.int16, device="cuda")
        ms = timed(lambda i: dec.decode(tk, raw_ids=True, seed=i, precision=args.precision, out=ok), 3 if F == 512 else 5, 1)
        sweep[str(F)] = {"ms": ms, "audio_s_per_s": world * 64 * F * 2048 / SR / (ms * 1e-3)}
        del tk, ok
    res["config3_b64_full_utterance_frames"] = sweep
    # config 4: Hindi-vocabulary streams (0.5 % out-of-range ids -> clamp), 512 streams per GPU, sliding-window call
    tk = torch.from_numpy(synth.make_tokens(512, FRAMES, seed=77 + rank, bad_frac=0.005)).cuda()
    ok = torch.empty((512, 2048), dtype=torch.int16, device="cuda")
    ms = timed(lambda i: dec.decode(tk, raw_ids=True, seed=i, extract_slice=True, precision=args.precision, out=ok), 10, 3)
    res["config4_hindi_vocab_512_streams_per_gpu"] = {
        "streams": 512 * world, "ms": ms, "windows_per_s": world * 512 / (ms * 1e-3),
        "emitted_audio_s_per_s": world * 512 * 2048 / SR / (ms * 1e-3), "bad_id_fraction": 0.005}
    return res


def neighbour_rows(dec, args):
    """The rows SURVEY.md section 8(f) puts either side of the path, measured on rank 0 (CUDA events): the stateful streaming
    session (row 1) against one batch decode of the same streams, and the SNAC encode path (row 4).  Beside the headline,
    never inside it."""
    import torch
    from tts_inference_b200 import synth
    from tts_inference_b200.encoder import SnacEncoder
    out = {}
    S, F_, k = 64, 64, 4
    tok = torch.from_numpy(synth.make_tokens(S, F_, seed=3)).cuda()
    keys = torch.arange(S, dtype=torch.int32).cuda()
    sess = dec.open_session(S, 32, precision=args.precision)

    def stream_once():
        sess.reset()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for f in range(0, F_, k):
            sess.step(0, tok[:, 7 * f:7 * (f + k)], final=(f + k == F_), seed=1, stream_keys=keys)
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b)
    stream_once()
    ms_s = min(stream_once() for _ in range(3))
    dec.decode(tok, raw_ids=True, seed=1, stream_keys=keys, precision=args.precision)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dec.decode(tok, raw_ids=True, seed=1, stream_keys=keys, precision=args.precision); b.record()
    torch.cuda.synchronize()
    audio_s = S * F_ * 2048 / SR
    out["streaming_session"] = {
        "what": f"{S} streams x {F_} frames fed {k} frames per step through snacb_session_step (32-frame sliding window per "
                "stream, per-stage state in HBM, only newly final rows computed; bit-identical to the batch decode)",
        "steps": F_ // k, "ms_all_steps": ms_s, "audio_s_per_s": audio_s / ms_s * 1e3,
        "one_batch_decode_ms": a.elapsed_time(b), "session_bytes": sess.nbytes}
    sess.close()
    # the reference's sliding 28/7 cadence (one new frame per stream per step) on the session: 1024 streams, 1 / 2 / 4 frames
    # per step, against the sliced window call that the sliding policy costs (sliding_window_mode above: one 4-frame window,
    # trimmed to the receptive field of its middle frame, per new frame)
    S2 = 1024
    tok2 = torch.from_numpy(synth.make_tokens(S2, 48, seed=5)).cuda()
    sess2 = dec.open_session(S2, 32, precision=args.precision)
    slots = np.arange(S2, dtype=np.int32)
    cad = {}
    for k2 in (1, 2, 4):
        # ASYNCHRONOUS streams: four cohorts that started 0 / 3 / 6 / 9 frames ago (so their windows also slide at
        # different steps); every step serves all 1024 in one launch sequence (snacb_session_step_multi)
        sess2.reset()
        pos = np.zeros(S2, dtype=np.int64)
        for c in range(4):
            sel = slots[c::4]
            f0 = 8 + 3 * (3 - c)
            sess2.step_multi(sel, tok2[c::4, :7 * f0].contiguous(), seed=1)
            pos[c::4] = f0
        evs = []
        for _ in range(16 // k2 + 2):
            new = torch.stack([tok2[i, 7 * int(pos[i]):7 * (int(pos[i]) + k2)] for i in range(4)])   # one row per cohort
            new = new.repeat(S2 // 4, 1).contiguous()            # cohort c = rows c, c + 4, ... (token VALUES do not matter here)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); sess2.step_multi(slots, new, seed=1); b.record()
            evs.append((a, b)); pos += k2
        torch.cuda.synchronize()
        ms = sorted(x.elapsed_time(y) for x, y in evs)[len(evs) // 2]
        cad[f"{k2}_frames_per_step"] = {"ms_per_step": ms, "emitted_audio_s_per_s": S2 * k2 * 2048 / SR / ms * 1e3}
    out["streaming_session"]["cadence_1024_async_streams"] = cad
    sess2.close()
    # the native streamer end to end: host token ids in, host int16 out, 1024 asynchronous streams, one new frame per stream
    # per tick (pushes outside the timed region: they run on producer threads in a server)
    from tts_inference_b200.streamer import StreamServer
    srv = StreamServer(dec, S2, 32, precision=args.precision, min_frames=1, max_samples_per_tick=S2 * 2048 * 12)
    host_tok = synth.make_tokens(S2, 40, seed=5)
    for i in range(S2):
        srv.push(i, host_tok[i, :7 * (8 + 3 * (i % 4))])
    srv.tick(seed=1)
    srv.tick(seed=1)                                         # (a first step takes at most window - 16 frames: the 17-frame cohort's rest)
    ticks = []
    pos = [8 + 3 * (i % 4) for i in range(S2)]
    for _ in range(10):
        for i in range(S2):
            srv.push(i, host_tok[i, 7 * pos[i]:7 * pos[i] + 7]); pos[i] += 1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        chunks = srv.tick(seed=1)
        ticks.append(time.perf_counter() - t0)
        assert len(chunks) == S2 and all(c.size == 2048 for _, c in chunks)
    tk = sorted(ticks)[len(ticks) // 2]
    out["streaming_session"]["native_streamer_tick_1024_async_streams"] = {
        "what": "snacb_streamer_tick wall time: staging + H2D of the new tokens, one session step for all streams, D2H of the PCM, sync",
        "ms_per_tick": tk * 1e3, "emitted_audio_s_per_s": S2 * 2048 / SR / tk}
    srv.close()
    enc = SnacEncoder(synth.make_encoder_state_dict(0))
    audio = torch.from_numpy(synth.make_audio(64, 2048 * 16)).cuda()
    enc.encode(audio)
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); enc.encode(audio); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out["encode_path"] = {"what": "SNAC encode (audio -> codes), 64 utterances x 16 frames, fp32 CUDA-core kernels",
                          "ms": min(ts), "audio_s_per_s": 64 * 16 * 2048 / SR / min(ts) * 1e3}
    enc.close()
    return out


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's own CPU implementation of the path = oracle port (the pip
    package `snac` cannot be installed offline), all host threads, bounded sample per step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = 16
    import torch
    from oracle import glue_ref, synth_ckpt
    from tts_inference_b200 import synth
    torch.set_num_threads(threads)
    model = synth_ckpt.make_model(0)
    codes = synth.make_tokens(batch, FRAMES).astype(np.int64) - 128266

    def step():
        lv = glue_ref.unpack_np(codes)
        y = model.decode([torch.from_numpy(x.astype(np.int64)) for x in lv])
        return glue_ref.pcm16_torch(y)

    for _ in range(max(args.warmup, 1)):
        step()
    t = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t) / args.steps
    v = batch * WINDOW_SAMPLES / SR / dt
    sample = f"{batch} windows x 28 tokens per step (of the arm's {args.batch}), oracle port, torch {torch.__version__} CPU fp32"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch), "sample_windows_per_step": batch},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(batch: int) -> str:
    return (f"batched streaming decode: {batch} concurrent streams x one 28-token (4-frame) window per step per GPU, "
            f"full window -> 8192 int16 samples (BASELINE configs[1] shape at the north_star batch; the sliced "
            f"[2048:4096] call is reported under sliding_window_mode)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip BASELINE configs 1/3/4 and the PyTorch-on-GPU baseline")
    ap.add_argument("--profile-out", default=None, help="write the per-stage table here (json)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from tts_inference_b200 import SnacDecoder, synth

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    B = args.batch
    sd = synth.make_state_dict(0)
    dec = SnacDecoder(sd, device=local_rank)
    # every rank decodes its own B streams (stream ids rank*B .. rank*B+B-1): weak scaling, no exchange
    tokens = synth.make_tokens(B, FRAMES, seed=20241224 + rank)
    tok_dev = torch.from_numpy(tokens).cuda()
    pcm_dev = torch.empty((B, WINDOW_SAMPLES), dtype=torch.int16, device="cuda")
    tok_pin = torch.from_numpy(tokens).pin_memory()
    pcm_pin = torch.empty((B, WINDOW_SAMPLES), dtype=torch.int16).pin_memory()
    pcm_dev_sl = torch.empty((B, 2048), dtype=torch.int16, device="cuda")
    pcm_pin_sl = torch.empty((B, 2048), dtype=torch.int16).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def step(i):
        dec.decode(tok_dev, raw_ids=True, extract_slice=False, seed=i, precision=args.precision, out=pcm_dev)

    def step_sliced(i):
        dec.decode(tok_dev, raw_ids=True, extract_slice=True, seed=i, precision=args.precision, out=pcm_dev_sl)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()

    # ------------------------------------------------------------------ timed region (device-resident inputs)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = dec.stats()[0]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                       # evict L2 between timed steps (not timed)
        ev[i][0].record()
        step(100 + i)
        ev[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = dec.stats()[0] - l0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop()

    # ------------------------------------------------------------------ e2e (host buffers, copies inside)
    # every step copies ITS tokens host->device and ITS PCM device->host (pinned buffers); the serving-loop form of the
    # call keeps one step in flight, so the copy-out of step i overlaps the decode of step i + 1
    # (snacb_decode_host_submit / _wait, include/snacb.h).  The blocking call is timed beside it.
    def e2e_loop(pins, sliced, seed0):
        for i in range(2):
            dec.decode_host_ptr(tok_pin.data_ptr(), B, 28, pins[0].data_ptr(), raw_ids=True, extract_slice=sliced,
                                seed=i, precision=args.precision)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            dec.decode_host_ptr(tok_pin.data_ptr(), B, 28, pins[0].data_ptr(), raw_ids=True, extract_slice=sliced,
                                seed=seed0 + i, precision=args.precision)
        blocking_s = time.perf_counter() - t0
        barrier()
        for i in range(2):                                   # warm-up of the pipelined form (allocates its two staging slots)
            dec.submit_host_ptr(tok_pin.data_ptr(), B, 28, pins[i & 1].data_ptr(), raw_ids=True, extract_slice=sliced,
                                seed=i, precision=args.precision)
        dec.wait_host(); dec.wait_host()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            dec.submit_host_ptr(tok_pin.data_ptr(), B, 28, pins[i & 1].data_ptr(), raw_ids=True, extract_slice=sliced,
                                seed=seed0 + 50 + i, precision=args.precision)
            if i:
                dec.wait_host()
        dec.wait_host()
        pipelined_s = time.perf_counter() - t0
        barrier()
        return blocking_s, pipelined_s

    pcm_pin2 = torch.empty((B, WINDOW_SAMPLES), dtype=torch.int16).pin_memory()
    e2e_blocking_s, e2e_s = e2e_loop((pcm_pin, pcm_pin2), False, 200)
    assert int((pcm_pin != 0).sum()) > B * 2048 and int((pcm_pin2 != 0).sum()) > B * 2048, "e2e output looks empty"

    # ------------------------------------------------------------------ sliding-window call (sliced, trimmed)
    for i in range(3):
        step_sliced(i)
    barrier()
    ev_sl = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()
        ev_sl[i][0].record()
        step_sliced(400 + i)
        ev_sl[i][1].record()
    barrier()
    sl_ms = sum(a.elapsed_time(b) for a, b in ev_sl)
    pcm_pin_sl2 = torch.empty((B, 2048), dtype=torch.int16).pin_memory()
    _, sl_e2e_s = e2e_loop((pcm_pin_sl, pcm_pin_sl2), True, 500)

    # max over ranks (tts_inference_b200.dist: all_reduce MAX, identity at N=1)
    from tts_inference_b200.dist import max_over_ranks
    dev_ms = max_over_ranks(dev_ms, device="cuda")
    e2e_ms = max_over_ranks(e2e_s * 1e3, device="cuda")
    e2e_blocking_s = max_over_ranks(e2e_blocking_s * 1e3, device="cuda") * 1e-3
    sl_ms = max_over_ranks(sl_ms, device="cuda")
    sl_e2e_ms = max_over_ranks(sl_e2e_s * 1e3, device="cuda")

    # ------------------------------------------------------------------ BASELINE configs 1 / 3 / 4 (all ranks)
    extra = None
    if not args.no_extra:
        extra = extra_configs(dec, args, world, max_over_ranks)

    # ------------------------------------------------------------------ per-stage profile (separate pass)
    prof = None
    if rank == 0:
        dec.profile(True)
        nprof = min(args.steps, 5)
        for i in range(nprof):
            step(300 + i)
        rep = dec.profile_report()
        dec.profile(False)
        tot = sum(ms for _, ms in rep.values())
        groups = max(1, rep["tail"][0] // nprof)
        S = -(-B // groups)
        fl = stage_flops(S)
        classes = {}
        for name, (cnt, ms) in rep.items():
            cls = name.split(".")[1].rstrip("012") if "." in name else name       # convt / noise / res / ...
            c = classes.setdefault(cls, {"ms": 0.0, "launches": 0, "flop": 0.0})
            c["ms"] += ms; c["launches"] += cnt; c["flop"] += fl.get(name, 0.0) * cnt
        top = max(classes, key=lambda k: classes[k]["ms"])
        prof = {"stages": {k: {"launches": c, "total_ms": ms, "share": ms / tot,
                               "tflops": (fl.get(k, 0.0) * c / (ms * 1e-3) / 1e12) if ms > 0 else 0.0}
                           for k, (c, ms) in rep.items()},
                "classes": {k: {"share": v["ms"] / tot, "avg_launch_ms": v["ms"] / v["launches"],
                                "tflops": v["flop"] / (v["ms"] * 1e-3) / 1e12} for k, v in classes.items()},
                "top": top, "streams_per_launch": S, "steps": nprof}
        if args.profile_out:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
            with open(args.profile_out, "w") as f:
                json.dump(prof, f, indent=1)

    # ------------------------------------------------------------------ latency mode (B=1 window), rank 0
    latency = None
    if rank == 0 and not args.no_latency:
        from tts_inference_b200.bench_util import measure_latency
        latency = measure_latency(200, args.precision, dec)

    if rank == 0:
        peaks = measured_peaks()
        windows = B * world * args.steps
        value = windows * WINDOW_SAMPLES / SR / (dev_ms * 1e-3)
        e2e = windows * WINDOW_SAMPLES / SR / (e2e_ms * 1e-3)
        topc = prof["classes"][prof["top"]]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")     # dram bytes per launch from the committed ncu capture
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(prof["top"], {}).get(str(B))
        roof = {"bound": "tensor", "kernel": {"res": "k_resunit2", "convt": "k_gemm_tc(convT)", "noise": "k_gemm_tc(noise)",
                                              "chain": "k_chain (NoiseBlock + 3 ResidualUnits fused)"}.get(prof["top"], prof["top"]),
                "achieved": topc["tflops"], "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": topc["tflops"] / peaks["bf16_tflops_sustained"], "traffic": traffic,
                "share_of_step": topc["share"], "avg_launch_ms": topc["avg_launch_ms"], "peak_source": peaks["source"],
                "note": "algorithmic FLOPs of the fused NoiseBlock + 3 ResidualUnits over the CUDA-event time of the k_chain launches; "
                        "the kernel is bound by the MUFU (XU) pipe and issue slots of its Snake / depthwise prologue, not by the "
                        "tensor pipe (ncu: XU 28-43 %, FMA 35-50 %, issue 46-64 %, tensor 7-18 %, DRAM 6-12 %; DESIGN.md sections 6.1, 6.2)",
                "whole_step": {"achieved": FLOP_PER_WINDOW * B * world * args.steps / (dev_ms * 1e-3) / 1e12 / world,
                               "frac": FLOP_PER_WINDOW * B * args.steps / (dev_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"]}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, dt = cpu_port_rate(16, 8, threads, state_dict=sd)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"8 x 16 windows (of {B}), oracle port of the reference PyTorch decoder, fp32, {dt * 1e3:.0f} ms per 16"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16 operands, f32 accumulate", "bf16": "bf16 operands, f32 accumulate", "fp32": "f32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": workload_name(B), "windows_per_gpu_per_step": B, "frames_per_window": FRAMES,
                       "weights": "random-init snac_24khz decode architecture (seed 0)",
                       "timing": "CUDA events per step, L2 flushed (256 MiB memset) between timed steps",
                       "windows_per_s": value * SR / WINDOW_SAMPLES,
                       "wall_s_timed_region": t_wall},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * 28 * 4, "d2h_bytes_per_step": B * WINDOW_SAMPLES * 2,
                    "call": "snacb_decode_host_submit / _wait (one step in flight: copy-out of step i overlaps decode of step i+1)",
                    "blocking_call_value": windows * WINDOW_SAMPLES / SR / max(e2e_blocking_s, 1e-9)},
            "sliding_window_mode": {
                "what": "same windows through the extract_slice=True call: samples [2048:4096] out, only their receptive "
                        "field computed (bit-identical to the full decode's slice; tests/test_gpu_parity.py::test_slice_semantics)",
                "windows_per_s": windows / (sl_ms * 1e-3), "ms_per_step": sl_ms / args.steps,
                "emitted_audio_s_per_s": windows * 2048 / SR / (sl_ms * 1e-3),
                "window_audio_s_per_s": windows * WINDOW_SAMPLES / SR / (sl_ms * 1e-3),
                "e2e_windows_per_s": windows / (sl_e2e_ms * 1e-3),
                "e2e_emitted_audio_s_per_s": windows * 2048 / SR / (sl_e2e_ms * 1e-3)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "latency_b1_window_ms": latency,
            "baseline_configs": extra,
            "gpu_torch_baseline": gpu_torch_baseline() if (world == 1 and not args.no_extra) else None,
            "neighbour_rows": neighbour_rows(dec, args) if not args.no_extra else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
