"""Per-stage CUDA-event times for one group of B windows (default 47), a few repetitions."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_inference_b200 import SnacDecoder, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 47
dec = SnacDecoder(synth.make_state_dict(0))
tok = torch.from_numpy(synth.make_tokens(B, 4)).cuda()
for i in range(3):
    dec.decode(tok, raw_ids=True, extract_slice=True, seed=i)
dec.profile(True)
for i in range(10):
    dec.decode(tok, raw_ids=True, extract_slice=True, seed=i)
rep = dec.profile_report()
tot = sum(ms for _, ms in rep.values())
print(" ".join(f"{k}={ms / c * 1e3:.1f}" for k, (c, ms) in rep.items()), f"total={tot / 10 * 1e3:.0f}us")
