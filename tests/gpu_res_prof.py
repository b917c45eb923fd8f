"""In-kernel wait / phase cycles of the block-0 ResidualUnit kernels (SNACB_RES_PROF=1): one decode of B windows."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SNACB_RES_PROF"] = "1"
from tts_inference_b200 import SnacDecoder, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dec = SnacDecoder(synth.make_state_dict(0))
tok = torch.from_numpy(synth.make_tokens(B, 4)).cuda()
for i in range(2):
    dec.decode(tok, raw_ids=True, seed=i)
torch.cuda.synchronize()
