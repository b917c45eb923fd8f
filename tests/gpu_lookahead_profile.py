"""Where the time of a ranged (lookahead-policy) decode call goes: CUDA-event time of the calls against the sum of their
kernels' own times (snacb_profile), per stage.   python tests/gpu_lookahead_profile.py [B] [F] [chunk]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_inference_b200 import SnacDecoder, policy, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
F = int(sys.argv[2]) if len(sys.argv) > 2 else 128
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dec = SnacDecoder(synth.make_state_dict(0))
tok = torch.from_numpy(synth.make_tokens(B, F, seed=1)).cuda()
keys = torch.arange(B, dtype=torch.int32).cuda()
sched, emitted = [], 0
for f in range(chunk, F + 1, chunk):
    end = policy.stable_samples(f, 5, False)
    if end > emitted:
        sched.append((f, emitted, end)); emitted = end
sched.append((F, emitted, 2048 * F))
toks = [tok[:, :7 * f].contiguous() for (f, _, _) in sched]


def run():
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for t, (f, lo, hi) in zip(toks, sched):
        dec.decode(t, raw_ids=True, seed=2, stream_keys=keys, sample_range=(lo, hi))
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


sess = dec.open_session(B, 32)
steps = list(range(chunk, F + 1, chunk)) + ([F] if F % chunk else [])
stoks = [tok[:, 7 * (f - chunk if f % chunk == 0 else f - f % chunk):7 * f].contiguous() for f in steps]


def run_stateful():
    sess.reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i, t in enumerate(stoks):
        sess.step(0, t, final=(i == len(stoks) - 1), seed=2, stream_keys=keys)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


if len(sys.argv) > 4 and sys.argv[4] == "stateful":
    run = run_stateful
run(); run()
ms = run()
dec.profile(True)
run()
rep = dec.profile_report()
dec.profile(False)
ksum = sum(v[1] for v in rep.values())
print(json.dumps({"calls": len(sched), "event_ms_all_calls_back_to_back": ms, "sum_of_kernel_ms": ksum,
                  "per_stage_ms": {k: round(v[1], 3) for k, v in rep.items()}}, indent=1))
