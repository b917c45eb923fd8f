"""Per-stage CUDA-event times (us per step) of one group of B windows, full and sliced call:
    python tests/gpu_stage_profile.py [B]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_inference_b200 import SnacDecoder, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dec = SnacDecoder(synth.make_state_dict(0))
tok = torch.from_numpy(synth.make_tokens(B, 4)).cuda()
for sl in (False, True):
    for i in range(3):
        dec.decode(tok, raw_ids=True, extract_slice=sl, seed=i)
    dec.profile(True)
    for i in range(5):
        dec.decode(tok, raw_ids=True, extract_slice=sl, seed=i)
    rep = dec.profile_report()
    dec.profile(False)
    tot = sum(ms for _, ms in rep.values())
    print(f"[sliced={sl}] total {tot / 5 * 1e3:.0f} us: " + " ".join(f"{k}={ms / c * 1e3:.0f}" for k, (c, ms) in rep.items()), flush=True)
