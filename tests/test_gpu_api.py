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


def _with_hooks(compat, noise, precision):
    class _H:
        def __enter__(self_):
            compat._test_noise, compat._test_precision = noise, precision
        def __exit__(self_, *a):
            compat._test_noise, compat._test_precision = None, None
    return _H()


def test_convert_to_audio_returns_the_reference_helpers_bytes(compat_mod):
    """The bytes the REFERENCE's own convert_to_audio returned (tests/golden/make_golden.py executes it out of
    /root/reference with the oracle as `snac`) against compat.convert_to_audio on the same codes and the same injected
    NoiseBlock noise, fp32 arithmetic: every glue case (ragged, out-of-range, specials, maximum codes), +-1 LSB."""
    z = np.load(os.path.join(GOLD, "decode_golden.npz"))
    with open(os.path.join(GOLD, "glue_golden.json")) as f:
        cases = json.load(f)["cases"]
    n = 0
    for c in cases:
        if "stream_pcm_sha256" not in c:
            assert compat_mod.convert_to_audio(c["codes"]) is None
            continue
        want = z["glue_pcm_" + c["name"]]
        t0 = 4 * (len(c["codes"]) // 7)
        noise = [torch.from_numpy(x).cuda() for x in synth.make_noises(1, t0)]
        with _with_hooks(compat_mod, noise, "fp32"):
            got = compat_mod.convert_to_audio(list(c["codes"]), False)
        assert isinstance(got, bytes) and len(got) == c["stream_pcm_len"], c["name"]
        d = np.abs(np.frombuffer(got, dtype=np.int16).astype(np.int32) - want.astype(np.int32))
        assert d.max() <= 1, (c["name"], int(d.max()))
        n += 1
    assert n >= 9


def test_compat_entry_points_against_decode_golden(compat_mod):
    """decode_golden.npz through the reference's entry-point NAMES: convert_to_audio (full and sliced) and
    redistribute_codes + decode_snac (pcm_trt), same noise, fp32: +-1 LSB; fp16 default path: SNR >= 40 dB."""
    z = np.load(os.path.join(GOLD, "decode_golden.npz"))
    tokens = z["tokens"]
    noises = synth.make_noises(tokens.shape[0], 16, seed=int(z["noise_seed"]))
    for b in range(tokens.shape[0]):
        codes = (tokens[b].astype(np.int64) - 128266).tolist()
        nb = [torch.from_numpy(np.ascontiguousarray(n[b:b + 1])).cuda() for n in noises]
        with _with_hooks(compat_mod, nb, "fp32"):
            full = np.frombuffer(compat_mod.convert_to_audio(codes, False), dtype=np.int16)
            sl = np.frombuffer(compat_mod.convert_to_audio(codes, True), dtype=np.int16)
            l0, l1, l2 = compat_mod.redistribute_codes(codes)
            trt = np.frombuffer(compat_mod.decode_snac(l0, l1, l2, compat_mod.snac_model, "cuda"), dtype=np.int16)
        for got, want in ((full, z["pcm_full"][b]), (sl, z["pcm_slice"][b]), (trt, z["pcm_trt"][b])):
            assert got.shape == want.shape
            assert np.abs(got.astype(np.int32) - want.astype(np.int32)).max() <= 1
        with _with_hooks(compat_mod, nb, "fp16"):
            h = np.frombuffer(compat_mod.convert_to_audio(codes, False), dtype=np.int16).astype(np.float64)
        w = z["pcm_full"][b].astype(np.float64)
        assert 10 * np.log10((w ** 2).sum() / ((w - h) ** 2).sum()) >= 40.0


def test_init_snac_from_checkpoint_file_and_env(tmp_path, monkeypatch, state_dict, oracle_model):
    """The real-checkpoint path: a pytorch_model.bin written with torch.save in BOTH weight-norm key styles is loaded by
    init_snac(path), by init_snac(directory) and by the reference's argument-less init_snac() through SNACB_CKPT; the
    decode equals the one from the in-memory state dict bit for bit, and the second load comes from the fold cache."""
    from tests._util import oracle_decode
    from tts_inference_b200 import compat, weights
    monkeypatch.setenv("SNACB_CACHE_DIR", str(tmp_path / "cache"))
    old_style = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in state_dict.items()}
    new_style = {}
    for k, v in old_style.items():
        k2 = k.replace(".weight_g", ".parametrizations.weight.original0").replace(".weight_v", ".parametrizations.weight.original1")
        new_style[k2] = v
    # keys a full snac_24khz checkpoint has and the decode path must ignore
    new_style["encoder.block.0.weight"] = torch.zeros(3)
    new_style["quantizer.quantizers.0.in_proj.bias"] = torch.zeros(8)
    d_old, d_new = tmp_path / "old", tmp_path / "snac_24khz"
    d_old.mkdir(); d_new.mkdir()
    torch.save(old_style, d_old / "pytorch_model.bin")
    torch.save(new_style, d_new / "pytorch_model.bin")
    tokens = synth.make_tokens(2, 4, seed=77)
    noises = synth.make_noises(2, 16, seed=4)
    nz = [torch.from_numpy(n).cuda() for n in noises]
    tok = torch.from_numpy(tokens).cuda()
    keep = (compat.snac_model, compat.snac_device)
    try:
        ref_dec = compat.init_snac(state_dict, device=0)
        want = ref_dec.decode(tok, raw_ids=True, noise=nz, precision="fp32", return_wave=True)
        outs = []
        outs.append(compat.init_snac(str(d_old / "pytorch_model.bin")))
        outs.append(compat.init_snac(str(d_new)))                     # directory, new-style keys, extra keys
        assert len(list((tmp_path / "cache").glob("folded-*.npz"))) == 2
        outs.append(compat.init_snac(str(d_new)))                     # cache hit
        monkeypatch.setenv("SNACB_CKPT", str(d_new))
        outs.append(compat.init_snac())                               # the reference's call
        for dec in outs:
            got = dec.decode(tok, raw_ids=True, noise=nz, precision="fp32", return_wave=True)
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        ref, _ = oracle_decode(oracle_model, tokens, noises)
        assert np.abs(want[1].cpu().numpy() - ref).max() <= 1e-3
        monkeypatch.delenv("SNACB_CKPT")
        with pytest.raises(RuntimeError, match="no checkpoint"):
            compat.init_snac()
        with pytest.raises(FileNotFoundError):
            compat.init_snac(str(tmp_path / "missing"))
    finally:
        compat.snac_model, compat.snac_device = keep


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


# ------------------------------------------------------------------ stateful streaming session
@pytest.mark.parametrize("precision,chunks", [("fp16", [4] * 6), ("fp16", [1, 3, 7, 2, 5, 6]), ("fp16", [24]), ("bf16", [4, 8, 12])])
def test_session_stream_is_bit_identical_to_batch_decode(decoder, precision, chunks):
    """Frames appended step by step to a session: each stage computes only the rows that became final, and the
    concatenated samples equal ONE keyed batch decode of the finished streams bit for bit."""
    F = sum(chunks)
    B = 3
    tokens = synth.make_tokens(B, F, seed=31)
    tok = torch.from_numpy(tokens).cuda()
    keys = torch.tensor([5, 11, 2], dtype=torch.int32).cuda()
    ref = decoder.decode(tok, raw_ids=True, seed=9, precision=precision, stream_keys=keys).cpu().numpy()
    sess = decoder.open_session(B + 2, F, precision=precision)
    assert sess.max_frames >= F and sess.nbytes > 0
    got, f0 = [], 0
    l0 = decoder.stats()[0]
    for i, k in enumerate(chunks):
        last = i == len(chunks) - 1
        if last:                                           # end of stream as a separate, token-less step
            got.append(sess.step(1, tok[:, 7 * f0: 7 * (f0 + k)], seed=9, stream_keys=keys).cpu().numpy())
            got.append(sess.step(1, tok[:, :0], final=True, seed=9, stream_keys=keys).cpu().numpy())
        else:
            got.append(sess.step(1, tok[:, 7 * f0: 7 * (f0 + k)], seed=9, stream_keys=keys).cpu().numpy())
        f0 += k
        assert sess.frames(1) == f0 and sess.frames(0) == 0
    cat = np.concatenate(got, axis=1)
    assert cat.shape == ref.shape
    assert np.array_equal(cat, ref)
    assert sess.emitted(2) == 2048 * F
    assert decoder.stats()[0] > l0
    with pytest.raises(Exception):                         # finished slots need a reset
        sess.step(1, tok[:, :7], seed=9)
    sess.reset(1, B)
    again = [sess.step(1, tok[:, :7 * F], final=True, seed=9, stream_keys=keys).cpu().numpy()]
    assert np.array_equal(again[0], ref)
    sess.close()


def test_session_default_keys_are_slot_indices_and_steps_cost_new_rows_only(decoder):
    """Without stream keys a slot's noise is keyed by its slot index whichever slots share its step; non-final steps emit
    what snacb_session_next_emit announces (the receptive-field lag, 2.5 frames); slots at different positions are refused."""
    F = 16
    tokens = synth.make_tokens(4, F, seed=4)
    tok = torch.from_numpy(tokens).cuda()
    ref = decoder.decode(tok, raw_ids=True, seed=1, stream_keys=torch.arange(4, dtype=torch.int32).cuda()).cpu().numpy()
    sess = decoder.open_session(4, F)
    outs = {s: [] for s in range(4)}
    # slots 0-1 run ahead of slots 2-3
    outs01 = sess.step(0, tok[:2, :7 * 8], seed=1).cpu().numpy()
    assert outs01.shape[1] == sess.emitted(0) and 2048 * 5.5 < outs01.shape[1] < 2048 * 6
    for s in (0, 1):
        outs[s].append(outs01[s])
    with pytest.raises(Exception):
        sess.step(0, tok[:, 7 * 8: 7 * 12], seed=1)                      # slots 0-1 hold 8 frames, slots 2-3 none
    o = sess.step(2, tok[2:, :7 * 8], seed=1).cpu().numpy()
    for s in (2, 3):
        outs[s].append(o[s - 2])
    n = sess.next_emit(0, 8, final=True)
    o = sess.step(0, tok[:, 7 * 8:], final=True, seed=1).cpu().numpy()      # all four together from here
    assert o.shape[1] == n
    for s in range(4):
        outs[s].append(o[s])
        assert np.array_equal(np.concatenate(outs[s]), ref[s]), s
    big = torch.from_numpy(synth.make_tokens(1, 40, seed=4)).cuda()
    small = decoder.open_session(1, 8)                                    # window of 32 frames (rounded up)
    with pytest.raises(Exception):
        small.step(0, big, seed=1)                                         # 40 frames at once do not fit the window
    small.step(0, big[:, :7 * 30], seed=1)
    with pytest.raises(Exception):
        small.step(0, big[:, :7 * 25], seed=1)                             # nor do 25 more: the window keeps 8 frames, 8 + 25 > 32
    with pytest.raises(ValueError):
        decoder.open_session(1, 8, precision="fp32")


@pytest.mark.parametrize("chunk", [4, 16, 7])
def test_session_window_slides_over_a_long_stream(decoder, chunk):
    """A stream much longer than the session's window (100 frames through 32): the window keeps its last 8 frames and
    moves them to the front whenever a step does not fit; the NoiseBlock noise stays keyed by the absolute time step.  The
    stream is still bit-identical to one batch decode, and the session's memory does not grow."""
    F, B = 100, 2
    tokens = synth.make_tokens(B, F, seed=77)
    tok = torch.from_numpy(tokens).cuda()
    keys = torch.tensor([3, 9], dtype=torch.int32).cuda()
    ref = decoder.decode(tok, raw_ids=True, seed=4, stream_keys=keys).cpu().numpy()
    sess = decoder.open_session(B, 32)
    nbytes = sess.nbytes
    got, f0 = [], 0
    while f0 < F:
        k = min(chunk, F - f0)
        got.append(sess.step(0, tok[:, 7 * f0: 7 * (f0 + k)], final=(f0 + k == F), seed=4, stream_keys=keys).cpu().numpy())
        f0 += k
        assert sess.frames(0) == f0
    cat = np.concatenate(got, axis=1)
    assert cat.shape == ref.shape and np.array_equal(cat, ref)
    assert sess.emitted(1) == 2048 * F and sess.nbytes == nbytes
    sess.close()


def test_session_c_abi_edge_cases(decoder):
    """Raw C-ABI calls: zero slots / zero frames are no-ops, bad ranges and null buffers are refused with a message, a
    padded pcm row stride is honoured, the first audio appears with the third frame (receptive-field lag of 2.5 frames;
    the reference's loops wait for four)."""
    import ctypes as C
    lib = decoder._lib
    sess = decoder.open_session(3, 32)
    s = sess._s
    got = C.c_int(-1)
    assert lib.snacb_session_step(s, 0, 0, None, 0, 0, 0, C.c_uint64(0), None, None, 0, C.byref(got), None) == 0 and got.value == 0
    assert lib.snacb_session_step(s, 2, 2, None, 0, 0, 0, C.c_uint64(0), None, None, 0, None, None) == -1      # slots 2..3 of 3
    assert lib.snacb_session_step(s, 0, 1, None, 7, 1, 0, C.c_uint64(0), None, None, 0, None, None) == -1      # null tokens
    assert b"null tokens" in lib.snacb_last_error(decoder._h)
    assert lib.snacb_session_reset(s, 1, 5) == -1 and lib.snacb_session_frames(s, 7) == -1
    tok = torch.from_numpy(synth.make_tokens(2, 6, seed=12)).cuda()
    ref = decoder.decode(tok, raw_ids=True, seed=5, stream_keys=torch.tensor([1, 2], dtype=torch.int32).cuda()).cpu().numpy()
    assert sess.next_emit(1, 2) == 0 and sess.next_emit(1, 3) == 2048 * 3 - 5050
    first = sess.step(1, tok[:, :14], seed=5)                           # two frames: nothing is final yet
    assert first.shape == (2, 0) and sess.frames(1) == 2 and sess.emitted(1) == 0
    assert lib.snacb_session_step(s, 1, 2, None, 0, 0, 0, C.c_uint64(5), None, None, 0, C.byref(got), None) == 0 and got.value == 0
    n = sess.next_emit(1, 1)
    wide = torch.full((2, n + 5), -7, dtype=torch.int16).cuda()          # row stride 5 samples wider than what is emitted
    rc = lib.snacb_session_step(s, 1, 2, tok[:, 14:21].contiguous().data_ptr(), 7, 1, 0, C.c_uint64(5), None, wide.data_ptr(),
                                n + 5, C.byref(got), None)
    assert rc == 0 and got.value == n == 2048 * 3 - 5050
    torch.cuda.synchronize()
    w = wide.cpu().numpy()
    assert np.array_equal(w[:, :n], ref[:, :n]) and (w[:, n:] == -7).all()
    small = torch.empty((2, 10), dtype=torch.int16).cuda()
    assert lib.snacb_session_step(s, 1, 2, tok[:, 21:].contiguous().data_ptr(), 21, 3, 1, C.c_uint64(5), None, small.data_ptr(),
                                  10, None, None) == -1                  # pcm rows too short for the flush
    rest = sess.step(1, tok[:, 21:], final=True, seed=5).cpu().numpy()
    assert np.array_equal(np.concatenate([w[:, :n], rest], axis=1), ref)
    sess.close()


@pytest.mark.parametrize("precision,k", [("fp16", 1), ("fp16", 3), ("bf16", 2)])
def test_session_step_multi_serves_asynchronous_streams(decoder, precision, k):
    """Streams that start at different times and therefore sit at different positions share ONE launch sequence per step
    (snacb_session_step_multi: per-stream row offset, buffer slot and window origin in every kernel), in scattered slots,
    with windows sliding at different times -- and each stream still equals its own batch decode bit for bit."""
    starts = [0, 0, 3, 7, 8, 20]                     # step at which a stream begins
    lens = [70, 41, 66, 30, 52, 12]                  # frames (several times the 32-frame window)
    lens = [n - n % k for n in lens]
    slots = [5, 0, 7, 2, 6, 3]                       # scattered, not contiguous, not ordered
    n_str = len(lens)
    tokens = synth.make_tokens(n_str, max(lens), seed=91)
    tok = torch.from_numpy(tokens).cuda()
    keys = torch.tensor([11, 4, 9, 1, 30, 7], dtype=torch.int32).cuda()
    refs = [decoder.decode(tok[i:i + 1, :7 * lens[i]].contiguous(), raw_ids=True, seed=6, precision=precision,
                           stream_keys=keys[i:i + 1]).cpu().numpy()[0] for i in range(n_str)]
    sess = decoder.open_session(8, 32, precision=precision)
    got = [[] for _ in lens]
    pos = [0] * n_str
    launches = []
    for step in range(200):
        live = [i for i in range(n_str) if step >= starts[i] and pos[i] < lens[i]]
        if not live and all(pos[i] >= lens[i] for i in range(n_str)):
            break
        # streams past their third frame share one call whatever their position; younger ones are grouped by exact position
        groups = {}
        for i in live:
            groups.setdefault(-1 if pos[i] >= 3 else pos[i], []).append(i)
        for _, members in sorted(groups.items()):
            new = torch.stack([tok[i, 7 * pos[i]: 7 * (pos[i] + k)] for i in members]).contiguous()
            l0 = decoder.stats()[0]
            out = sess.step_multi([slots[i] for i in members], new, seed=6, stream_keys=keys[members].contiguous()).cpu().numpy()
            launches.append((len(members), decoder.stats()[0] - l0))
            for row, i in enumerate(members):
                got[i].append(out[row])
                pos[i] += k
        for i in range(n_str):
            if pos[i] == lens[i] and sess.frames(slots[i]) == lens[i] and sess.emitted(slots[i]) < 2048 * lens[i]:
                got[i].append(sess.step(slots[i], tok[i:i + 1, :0], final=True, seed=6, stream_keys=keys[i:i + 1]).cpu().numpy()[0])
    for i in range(n_str):
        cat = np.concatenate(got[i])
        assert cat.shape == refs[i].shape, (i, cat.shape)
        assert np.array_equal(cat, refs[i]), i
    # a step costs the same number of launches whether it serves one stream or five
    assert len({l for _, l in launches if l > 0}) <= 3, sorted(set(launches))
    assert max(m for m, _ in launches) >= 4
    with pytest.raises(Exception):
        sess.step_multi([1, 1], tok[:2, :7].contiguous())                 # repeated slot
    sess.close()


def test_native_streamer_from_producer_threads(decoder):
    """snacb_streamer_*: four producer threads push token ids of 12 streams (different lengths, starting at different
    times, more streams over time than there are slots) while the main thread ticks; every stream's chunks concatenate to
    its own batch decode, bit for bit; slots are recycled; a push beyond the slot capacity is refused."""
    import threading
    import time
    from tts_inference_b200.streamer import StreamServer
    n_str, max_streams = 12, 8
    lens = [9, 40, 23, 5, 61, 17, 33, 2, 28, 45, 12, 36]
    tokens = synth.make_tokens(n_str, max(lens), seed=123)
    srv = StreamServer(decoder, max_streams, 32, min_frames=1)
    got = {i: [] for i in range(n_str)}
    order = []                                   # stream ids in the order of their first push = their noise keys
    lock = threading.Lock()
    stop = threading.Event()
    errors = []

    def producer(my):
        try:
            for sid in my:
                while True:                      # wait for a free slot (the first push claims it)
                    try:
                        with lock:
                            srv.push(sid, tokens[sid, :3].tolist())
                            order.append(sid)
                        break
                    except Exception:
                        time.sleep(0.002)
                pos = 3
                n_ids = 7 * lens[sid] + (sid % 4)            # some streams end with a ragged tail (< 7 ids): dropped
                ids = np.concatenate([tokens[sid, :7 * lens[sid]], tokens[sid, :sid % 4]])
                while pos < n_ids:
                    step = 1 + (sid + pos) % 11
                    srv.push(sid, ids[pos:pos + step].tolist())
                    pos += step
                    if pos % 3 == 0:
                        time.sleep(0.0005)
                srv.end(sid)
        except Exception as e:                   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=producer, args=(list(range(k, n_str, 4)),)) for k in range(4)]
    for t in threads:
        t.start()
    t0 = time.time()
    while (any(t.is_alive() for t in threads) or srv.active() > 0) and time.time() - t0 < 120:
        for sid, pcm in srv.tick(seed=4):
            got[sid].append(pcm)
    for t in threads:
        t.join()
    assert not errors, errors
    assert srv.active() == 0
    for sid in range(n_str):
        key = torch.tensor([order.index(sid)], dtype=torch.int32).cuda()
        tok = torch.from_numpy(np.ascontiguousarray(tokens[sid:sid + 1, :7 * lens[sid]])).cuda()
        ref = decoder.decode(tok, raw_ids=True, seed=4, stream_keys=key).cpu().numpy()[0]
        cat = np.concatenate(got[sid]) if got[sid] else np.empty(0, dtype=np.int16)
        assert cat.shape == ref.shape, (sid, cat.shape, ref.shape)
        assert np.array_equal(cat, ref), sid
    # capacity: the ninth concurrent stream is refused until a slot frees up
    for sid in range(100, 100 + max_streams):
        srv.push(sid, tokens[0, :7].tolist())
    with pytest.raises(Exception):
        srv.push(999, tokens[0, :7].tolist())
    srv.close()


def test_two_gpus_in_one_process(state_dict):
    """Two handles on two GPUs in ONE process (the opt-in shared-memory attributes are per device; entry points
    must select their handle's device): the same tokens decode to the same bytes on both, calls interleaved, a session and
    the device ingest on the second GPU while the current device is the first."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from tts_inference_b200 import SnacDecoder
    from tts_inference_b200.ingest import DeviceIngest
    d0, d1 = SnacDecoder(state_dict, device=0), SnacDecoder(state_dict, device=1)
    tokens = synth.make_tokens(33, 4, seed=8)
    t0, t1 = torch.from_numpy(tokens).cuda(0), torch.from_numpy(tokens).cuda(1)
    torch.cuda.set_device(0)
    outs = []
    for i in range(3):
        a = d0.decode(t0, raw_ids=True, seed=i, extract_slice=bool(i & 1))
        b = d1.decode(t1, raw_ids=True, seed=i, extract_slice=bool(i & 1))        # current device is 0: the handle selects 1
        outs.append((a, b))
    for a, b in outs:
        assert a.device.index == 0 and b.device.index == 1
        assert torch.equal(a.cpu(), b.cpu())
    h = d1.decode_host(tokens, raw_ids=True, seed=0)
    assert np.array_equal(h, outs[0][0].cpu().numpy())
    sess = d1.open_session(4, 32)
    long_tok = torch.from_numpy(synth.make_tokens(2, 12, seed=3)).cuda(1)
    ref = d0.decode(long_tok.cuda(0), raw_ids=True, seed=2, stream_keys=torch.tensor([0, 1], dtype=torch.int32).cuda(0)).cpu()
    got = torch.cat([sess.step(0, long_tok[:, :42], seed=2).cpu(), sess.step(0, long_tok[:, 42:], final=True, seed=2).cpu()], dim=1)
    assert torch.equal(got, ref)
    ing = DeviceIngest(8, device=1) if "device" in DeviceIngest.__init__.__code__.co_varnames else None
    if ing is not None:
        ids = torch.full((8, 30), 128266 + 5, dtype=torch.int32).cuda(1)
        ids[:, 0] = 128257
        wt, ws, *_ = ing.step(ids)
        assert wt.shape[0] == 8 and wt.device.index == 1
    d0.close(); d1.close()                                  # closes d1's open session first
    assert not sess._s.value
