"""A small pass over every kernel of the library, meant to be run under compute-sanitizer on a GPU box:

    compute-sanitizer --tool memcheck python tests/gpu_sanitize.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_inference_b200 import SnacDecoder, egress, synth  # noqa: E402
from tts_inference_b200.ingest import DeviceIngest  # noqa: E402

dec = SnacDecoder(synth.make_state_dict(0))
tok = torch.from_numpy(synth.make_tokens(40, 4, bad_frac=0.02)).cuda()
for prec in ("fp16", "bf16", "fp32"):
    full = dec.decode(tok, raw_ids=True, seed=1, precision=prec)
    sl = dec.decode(tok, raw_ids=True, seed=1, precision=prec, extract_slice=True)
    assert torch.equal(sl, full[:, 2048:4096]), prec
long_tok = torch.from_numpy(synth.make_tokens(3, 24, seed=2)).cuda()
keys = torch.tensor([7, 8, 9], dtype=torch.int32).cuda()
a = dec.decode(long_tok, raw_ids=True, seed=3, stream_keys=keys)
b = dec.decode(long_tok, raw_ids=True, seed=3, stream_keys=keys, sample_range=(30000, 36000))
assert torch.equal(b, a[:, 30000:36000])
u = dec.decode(tok[:5], raw_ids=True, seed=1, unfused=True)
tp, pp = torch.from_numpy(synth.make_tokens(6, 4)).pin_memory(), torch.zeros((6, 8192), dtype=torch.int16).pin_memory()
dec.submit_host_ptr(tp.data_ptr(), 6, 28, pp.data_ptr(), raw_ids=True, seed=4)
dec.wait_host()
ing = DeviceIngest(50)
ids = torch.full((50, 40), 128266 + 5, dtype=torch.int32).cuda()
ids[:, 0] = 128257
ids[::3, 37] = 128258
w = ing.step(ids, finish=torch.ones(50, dtype=torch.uint8).cuda())
pcm = torch.randint(-32768, 32767, (9, 4097), dtype=torch.int16).cuda()
b64, wav = egress.pcm_to_base64(pcm), egress.pcm_to_wav(pcm)
torch.cuda.synchronize()
print("sanitize pass ok", tuple(full.shape), int(w[0].shape[0]), int(w[2].shape[0]), tuple(b64.shape), tuple(wav.shape), int(np.abs(u.cpu().numpy()).max()))
