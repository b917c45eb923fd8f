"""Cost of the lookahead streaming policy (tts_inference_b200/policy.py; tensorrt_tts/PIPELINE_REPORT.md:475-511) for B
streams of F frames, a decode every `chunk` new frames, 5 frames of lookahead -- CUDA-event time of all decode calls:
  (a) the reference's algorithm: re-decode ALL frames every time;   (b) snacb_decode_range: new stable samples only;
  (c) the stateful session (snacb_session_step): per-stage state in HBM, only newly final rows are computed;
  (d) one batch decode of the finished utterances (lower bound).

    python tests/gpu_lookahead_bench.py [B] [F] [chunk] > gpurun_out/lookahead_bench.json
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_inference_b200 import SnacDecoder, policy, synth  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    F = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    dec = SnacDecoder(synth.make_state_dict(0))
    tok = torch.from_numpy(synth.make_tokens(B, F, seed=1)).cuda()
    keys = torch.arange(B, dtype=torch.int32).cuda()
    sched = []
    emitted = 0
    for f in list(range(chunk, F + 1, chunk)) + ([F] if F % chunk else []):
        end = policy.stable_samples(f, 5, False)
        if end > emitted:
            sched.append((f, emitted, end)); emitted = end
    sched.append((F, emitted, 2048 * F))

    def run(mode):
        ms = 0.0
        outs = []
        for (f, lo, hi) in sched:
            t = tok[:, :7 * f].contiguous()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if mode == "full":
                o = dec.decode(t, raw_ids=True, seed=2, stream_keys=keys)[:, lo:hi]
            else:
                o = dec.decode(t, raw_ids=True, seed=2, stream_keys=keys, sample_range=(lo, hi))
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
            outs.append(o)
        return ms, torch.cat(outs, dim=1)

    sess = dec.open_session(B, 32)                                # 32-frame sliding window per stream, whatever F is

    def run_stateful():
        """the stateful session: every step appends `chunk` frames and emits what became final; no prefix is re-read"""
        sess.reset()
        ms, outs, f0 = 0.0, [], 0
        steps = list(range(chunk, F + 1, chunk)) + ([F] if F % chunk else [])
        for i, f in enumerate(steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t = tok[:, 7 * f0:7 * f].contiguous()
            a.record()
            o = sess.step(0, t, final=(i == len(steps) - 1), seed=2, stream_keys=keys)
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
            outs.append(o); f0 = f
        return ms, torch.cat(outs, dim=1), len(steps)

    run("range"); run("full"); run_stateful()                   # warm-up (workspace growth)
    ms_state, pcm_state, n_state = run_stateful()
    ms_full, pcm_full = run("full")
    ms_range, pcm_range = run("range")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dec.decode(tok, raw_ids=True, seed=2, stream_keys=keys)
    a.record(); batch = dec.decode(tok, raw_ids=True, seed=2, stream_keys=keys); b.record()
    torch.cuda.synchronize()
    ms_batch = a.elapsed_time(b)
    audio_s = B * F * 2048 / 24000.0
    print(json.dumps({
        "streams": B, "frames": F, "frames_per_chunk": chunk, "lookahead_frames": 5, "decode_calls": len(sched),
        "redecode_all_ms": ms_full, "ranged_ms": ms_range, "stateful_session_ms": ms_state, "one_batch_decode_ms": ms_batch,
        "stateful_session_steps": n_state, "stateful_session_bytes": sess.nbytes,
        "audio_s_per_s": {"redecode_all": audio_s / ms_full * 1e3, "ranged": audio_s / ms_range * 1e3,
                          "stateful_session": audio_s / ms_state * 1e3, "one_batch_decode": audio_s / ms_batch * 1e3},
        "streamed_equals_batch_decode": bool(torch.equal(pcm_range, batch) and torch.equal(pcm_full, batch)
                                             and torch.equal(pcm_state, batch)),
    }, indent=1))


if __name__ == "__main__":
    main()
