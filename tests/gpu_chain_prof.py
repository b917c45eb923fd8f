"""In-kernel phase timing of the chain kernels (SNACB_CHAIN_PROF=1): one decode of B windows, stderr carries the report."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SNACB_CHAIN_PROF"] = "1"
from tts_inference_b200 import SnacDecoder, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sl = len(sys.argv) > 2 and sys.argv[2] == "sliced"
dec = SnacDecoder(synth.make_state_dict(0))
tok = torch.from_numpy(synth.make_tokens(B, 4)).cuda()
for i in range(2):
    dec.decode(tok, raw_ids=True, seed=i, extract_slice=sl)
torch.cuda.synchronize()
