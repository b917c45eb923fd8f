import os, sys, torch
sys.path.insert(0, '/root/repo')
from tts_inference_b200 import SnacDecoder, synth
dec = SnacDecoder(synth.make_state_dict(0))
tok = torch.from_numpy(synth.make_tokens(1024, 4)).cuda()
for i in range(2):
    out = dec.decode(tok, raw_ids=True, seed=i)
torch.cuda.synchronize()
print("ok", out.shape)
