"""GPU tests of the reference-facing surface: the compat helper names and the window batcher."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import glue_ref
from tts_inference_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def compat_mod(state_dict):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tts_inference_b200 import compat
    compat.init_snac(state_dict, device=0)
    return compat


def test_convert_to_audio_contract(compat_mod):
    codes = synth.make_codes(1, 6)[0].tolist()
    assert compat_mod.convert_to_audio(codes[:6]) is None                 # modal_audio_stream.py:153-154
    full = compat_mod.convert_to_audio(codes[:28] + [1, 2, 3], extract_slice=False)
    assert isinstance(full, bytes) and len(full) == 2 * 8192             # ragged tail dropped
    sl = compat_mod.convert_to_audio(codes[:28], extract_slice=True)
    assert len(sl) == 2 * 2048
    one = compat_mod.convert_to_audio(codes[:7], extract_slice=True)      # 2048 samples: not > 4096, all kept
    assert len(one) == 2 * 2048
    huge = [10 ** 12, -10 ** 12] + codes[2:28]                            # Python ints beyond int32 clamp like any bad id
    assert len(compat_mod.convert_to_audio(huge)) == 2 * 8192


def test_redistribute_codes_matches_reference_vectors(compat_mod):
    with open(os.path.join(GOLD, "glue_golden.json")) as f:
        cases = json.load(f)["cases"]
    n = 0
    for c in cases:
        if "trt_levels" not in c:
            continue
        got = compat_mod.redistribute_codes(c["codes"])
        assert [list(x) for x in got] == c["trt_levels"], c["name"]
        n += 1
    assert n >= 8
    assert compat_mod.redistribute_codes([1, 2, 3]) == ([], [], [])


def test_decode_snac_equals_convert_to_audio_shapes(compat_mod):
    codes = synth.make_codes(1, 4, seed=9)[0].tolist()
    l0, l1, l2 = compat_mod.redistribute_codes(codes)
    pcm = compat_mod.decode_snac(l0, l1, l2, compat_mod.snac_model, compat_mod.snac_device)
    assert len(pcm) == 2 * 8192
    a = np.frombuffer(pcm, dtype=np.int16).astype(np.float64)
    b = np.frombuffer(compat_mod.convert_to_audio(codes), dtype=np.int16).astype(np.float64)
    # different noise draws (the reference draws fresh randn per decode): highly correlated, not equal
    assert np.corrcoef(a, b)[0, 1] > 0.7


def test_batcher_chunk_policy_matches_stream_audio(decoder):
    """policy 0 == stream_audio's buffer rule (modal_audio_stream.py:352-396): 28-code chunks, then the
    remaining whole frames at end of stream -- and equals decoding those chunks directly."""
    from tts_inference_b200.batcher import POLICY_CHUNK, WindowBatcher
    b = WindowBatcher(decoder, policy=POLICY_CHUNK, raw_ids=True, max_windows=64)
    streams = {sid: synth.make_tokens(1, 11, seed=sid)[0][: 7 * 11 - 3 * sid] for sid in (1, 2, 3)}
    # interleaved pushes of uneven sizes
    pos = {sid: 0 for sid in streams}
    step = {1: 5, 2: 13, 3: 28}
    while any(pos[s] < len(streams[s]) for s in streams):
        for sid in streams:
            if pos[sid] < len(streams[sid]):
                b.push(sid, streams[sid][pos[sid]: pos[sid] + step[sid]])
                pos[sid] += step[sid]
    for sid in streams:
        b.end(sid)
    out = b.flush(seed=5)
    got = {sid: [c for s, c in out if s == sid] for sid in streams}
    for sid, toks in streams.items():
        want = glue_ref.stream_chunks(toks.tolist())
        assert [len(c) for c in got[sid]] == [2048 * (len(w) // 7) for w in want]
    # the same full windows decoded directly (batch rows in flush order, same seed) give the same PCM
    full_rows = [(s, i) for i, (s, c) in enumerate(out) if len(c) == 8192]
    toks = []
    seen = {sid: 0 for sid in streams}
    for s, _ in full_rows:
        k = seen[s]; seen[s] += 1
        toks.append(streams[s][28 * k: 28 * k + 28])
    direct = decoder.decode_host(np.stack(toks), raw_ids=True, seed=5)
    for row, (s, i) in enumerate(full_rows):
        assert np.array_equal(direct[row], out[i][1])
    assert b.pending() == 0 and b.flush() == []


def test_batcher_sliding_policy(decoder):
    from tts_inference_b200.batcher import POLICY_SLIDING, WindowBatcher
    b = WindowBatcher(decoder, policy=POLICY_SLIDING, raw_ids=True, max_windows=64)
    toks = synth.make_tokens(1, 8, seed=3)[0]
    for i in range(0, len(toks), 3):
        b.push(42, toks[i:i + 3])
    out = b.flush(seed=1)
    wins = glue_ref.sliding_windows(toks.tolist())
    assert len(out) == len(wins) == 5 and all(s == 42 and len(c) == 2048 for s, c in out)
    direct = decoder.decode_host(np.asarray(wins, dtype=np.int32), raw_ids=True, extract_slice=True, seed=1)
    for i, (_, c) in enumerate(out):
        assert np.array_equal(c, direct[i])


def test_cuda_graph_replay_is_bit_identical(decoder):
    tok = torch.from_numpy(synth.make_tokens(1, 4, seed=2)).cuda()
    out = torch.empty((1, 2048), dtype=torch.int16, device="cuda")
    ref = decoder.decode(tok, raw_ids=True, extract_slice=True, seed=7).clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        decoder.decode(tok, raw_ids=True, extract_slice=True, seed=7, out=out)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        decoder.decode(tok, raw_ids=True, extract_slice=True, seed=7, out=out)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
