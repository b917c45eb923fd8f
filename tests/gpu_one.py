"""One decode of B windows (default 48) in the given precision -- the short command profiled under ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_inference_b200 import SnacDecoder, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dec = SnacDecoder(synth.make_state_dict(0))
tok = torch.from_numpy(synth.make_tokens(B, 4)).cuda()
for i in range(reps):
    out = dec.decode(tok, raw_ids=True, extract_slice=True, seed=i, precision=prec)
torch.cuda.synchronize()
print("ok", out.shape, int((out != 0).sum()), dec.stats())
