"""Token ingest and egress formats.

CPU: the oracle (oracle/ingest_ref.py) against the golden vectors produced by the reference's own
``generate_audio_tokens`` + ``stream_audio`` (tests/golden/make_golden_ingest.py) and against itself step-wise.
GPU: ``k_ingest`` / ``k_base64`` / ``k_wav`` through the C ABI, bit-exact against the oracle / the stdlib.
"""
import json
import os

import numpy as np
import pytest

from oracle import ingest_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "ingest_golden.json")))["cases"]
for _c in GOLDEN:     # what the engine delivered: vLLM finishes a request with its stop token (stop_token_ids=[TOKEN_EOS], :295)
    _c["delivered"] = _c["ids"][:_c["tokens_consumed"]]
SOS, EOS, BASE = ingest_ref.TOKEN_SOS, ingest_ref.TOKEN_EOS, ingest_ref.TOKEN_AUDIO_BASE


# ------------------------------------------------------------------------------------------------ CPU
def test_oracle_matches_reference_stream_audio():
    assert len(GOLDEN) >= 30 and {c["kind"] for c in GOLDEN} >= {"plain", "no_eos", "no_sos", "double_sos", "specials"}
    for c in GOLDEN:
        assert ingest_ref.ingest_stream(c["delivered"]) == c["chunks"], c["kind"]
        if c["kind"] != "eos_first":          # the filter itself also stops at TOKEN_EOS: ids after it change nothing
            assert ingest_ref.ingest_stream(c["ids"]) == c["chunks"], c["kind"]
        assert ingest_ref.last_sos_audio_tokens(c["ids"]) == c["last_sos_audio_tokens"], c["kind"]


@pytest.mark.parametrize("step", [1, 3, 7, 28, 64])
def test_stepwise_oracle_equals_whole_stream(step):
    """Feeding the streams a few ids per LLM step (the device kernel's shape) gives the same chunks."""
    streams = [c["delivered"] for c in GOLDEN]
    states = [ingest_ref.StreamState() for _ in streams]
    got = [[] for _ in streams]
    n = max(len(s) for s in streams)
    for t0 in range(0, n + step, step):
        toks = [s[t0:t0 + step] for s in streams]
        fin = [t0 + step >= len(s) for s in streams]          # generator exhausted (max_tokens): flush
        full, tails = ingest_ref.ingest_steps(states, toks, fin)
        for s, ids in full + tails:
            got[s].append([i - BASE for i in ids])
    # per stream the remainder always follows the full windows, so appending tails after fulls keeps time order
    for c, g in zip(GOLDEN, got):
        assert g == c["chunks"], c["kind"]


def test_egress_oracle_is_the_stdlib():
    pcm = np.arange(-5, 6, dtype=np.int16).tobytes()
    assert ingest_ref.b64(pcm[:7]).endswith(b"=") and len(ingest_ref.b64(pcm)) == 4 * ((len(pcm) + 2) // 3)
    w = ingest_ref.wav_bytes(pcm)
    assert w[:4] == b"RIFF" and w[8:16] == b"WAVEfmt " and len(w) == 44 + len(pcm) and w[44:] == pcm


# ------------------------------------------------------------------------------------------------ GPU
def _run_device(ing, streams, step, torch):
    S = len(streams)
    got = [[] for _ in range(S)]
    n = max(len(s) for s in streams)
    for t0 in range(0, n + step, step):
        tok = np.zeros((S, step), dtype=np.int32)
        nv = np.zeros(S, dtype=np.int32)
        for s, ids in enumerate(streams):
            part = ids[t0:t0 + step]
            tok[s, :len(part)] = part
            nv[s] = len(part)
        fin = np.array([t0 + step >= len(s) for s in streams], dtype=np.uint8)
        wt, ws, tt, ts, tf = ing.step(torch.from_numpy(tok).cuda(), torch.from_numpy(nv).cuda(), torch.from_numpy(fin).cuda())
        wt, ws, tt, ts, tf = (x.cpu().numpy() for x in (wt, ws, tt, ts, tf))
        assert list(ws) == sorted(ws) and list(ts) == sorted(ts)          # (stream, time) order
        for i, s in enumerate(ws):
            got[s].append([int(v) - BASE for v in wt[i]])
        for i, s in enumerate(ts):
            assert (tt[i, 7 * tf[i]:] == 0).all()
            got[s].append([int(v) - BASE for v in tt[i, :7 * tf[i]]])
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("step", [1, 5, 28, 97])
def test_device_ingest_matches_reference_chunks(step):
    import torch
    from tts_inference_b200.ingest import DeviceIngest
    streams = [c["delivered"] for c in GOLDEN]
    ing = DeviceIngest(len(streams))
    got = _run_device(ing, streams, step, torch)
    for c, g in zip(GOLDEN, got):
        assert g == c["chunks"], (c["kind"], step)
    st, cnt = ing.state()
    assert (st == 2).all() and (cnt == 0).all()
    # slots are reusable after a reset
    ing.reset()
    got = _run_device(ing, streams[::-1], 13, torch)
    for c, g in zip(GOLDEN[::-1], got):
        assert g == c["chunks"]


@pytest.mark.gpu
def test_device_ingest_many_streams_against_oracle():
    """3000 streams (three rounds of the single CTA), random step sizes, ids with specials; bit-exact vs the oracle."""
    import torch
    from tts_inference_b200.ingest import DeviceIngest
    rng = np.random.default_rng(3)
    S = 3000
    streams = []
    for s in range(S):
        n = int(rng.integers(0, 120))
        ids = (BASE + rng.integers(-20, 28672 + 20, size=n)).tolist()
        pre = rng.integers(0, 128000, size=int(rng.integers(0, 4))).tolist()
        if rng.random() < 0.9:
            pre.append(SOS)
        if rng.random() < 0.7:
            ids.append(EOS)
            ids += (BASE + rng.integers(0, 4096, size=int(rng.integers(0, 5)))).tolist()
        streams.append([int(v) for v in pre + ids])
    ing = DeviceIngest(S)
    got = _run_device(ing, streams, 9, torch)
    for s in range(S):
        assert got[s] == ingest_ref.ingest_stream(streams[s]), s


@pytest.mark.gpu
def test_ingest_then_decode_equals_decode_of_reference_chunks(decoder):
    """The device path end to end, step by step: ids -> windows -> PCM is bit-identical to decoding, with the same
    seed and batch order, the chunks the oracle's restatement of the reference loop forms at that step; and over the
    whole run every stream gets exactly the chunks of the reference's own stream_audio (golden)."""
    import torch
    from tts_inference_b200.ingest import DeviceIngest
    cases = GOLDEN[:12]
    streams = [c["delivered"] for c in cases]
    ing = DeviceIngest(len(streams))
    states = [ingest_ref.StreamState() for _ in streams]
    n_chunks = [0] * len(streams)
    step = 11
    n = max(len(s) for s in streams)

    def dec(rows):
        return decoder.decode(torch.tensor(rows, dtype=torch.int32).cuda(), raw_ids=True, seed=5).cpu().numpy()

    for t0 in range(0, n + step, step):
        parts = [ids[t0:t0 + step] for ids in streams]
        tok = np.zeros((len(streams), step), dtype=np.int32)
        for s, part in enumerate(parts):
            tok[s, :len(part)] = part
        nv = np.array([len(p) for p in parts], dtype=np.int32)
        fin = np.array([t0 + step >= len(s) for s in streams], dtype=np.uint8)
        got = ing.step_decode(decoder, torch.from_numpy(tok).cuda(), torch.from_numpy(nv).cuda(),
                              torch.from_numpy(fin).cuda(), seed=5)
        full, tails = ingest_ref.ingest_steps(states, parts, fin)
        want = []
        if full:
            pcm = dec([ids for _, ids in full])
            want += [(s, i, pcm[i]) for i, (s, _) in enumerate(full)]
        for fr in sorted({len(ids) // 7 for _, ids in tails}):
            grp = [(s, ids) for s, ids in tails if len(ids) // 7 == fr]
            pcm = dec([ids for _, ids in grp])
            want += [(s, 1 << 30, pcm[j]) for j, (s, _) in enumerate(grp)]
        want.sort(key=lambda e: (e[0], e[1]))
        assert [s for s, _ in got] == [s for s, _, _ in want]
        for (s, p), (_, _, w) in zip(got, want):
            assert np.array_equal(p.cpu().numpy(), w)
            n_chunks[s] += 1
    assert n_chunks == [len(c["chunks"]) for c in cases]


@pytest.mark.gpu
@pytest.mark.parametrize("n,samples", [(1, 2048), (3, 8192), (5, 1), (4, 2), (2, 4097), (1024, 8192), (7, 0)])
def test_base64_and_wav_are_byte_exact(n, samples):
    import torch
    from tts_inference_b200 import egress
    rng = np.random.default_rng(n * 10007 + samples)
    pcm = rng.integers(-32768, 32768, size=(n, samples), dtype=np.int16)
    d = torch.from_numpy(pcm).cuda()
    b = egress.pcm_to_base64(d).cpu().numpy()
    w = egress.pcm_to_wav(d).cpu().numpy()
    rows = range(n) if n <= 8 else (0, 1, n // 2, n - 1)
    for i in rows:
        assert b[i].tobytes() == ingest_ref.b64(pcm[i].tobytes())
        assert w[i].tobytes() == ingest_ref.wav_bytes(pcm[i].tobytes())


@pytest.mark.gpu
def test_io_bad_arguments():
    import ctypes as C
    from tts_inference_b200 import _lib
    lib = _lib.load()
    g = C.c_void_p()
    assert lib.snacb_ingest_create(C.byref(g), 0, 0) == -1
    assert lib.snacb_ingest_create(C.byref(g), 0, 4) == 0
    assert lib.snacb_ingest_reset(g, 2, 3, None) == -1
    assert lib.snacb_ingest_step(g, None, 4, 1, None, None, None, None, 0, None, None, None, None, None) == -1
    lib.snacb_ingest_destroy(g)
    assert lib.snacb_pcm_to_wav(None, 1, 4, 0, None, None) == -1
    assert lib.snacb_base64_len(5) == 8 and lib.snacb_ingest_window_capacity(10, 29) == 30


# ------------------------------------------------------------------------------------------------ lookahead policy
class _FakeDecoder:
    """Stands in for SnacDecoder on the CPU: sample t of a row is frames * 1_000_000 + t, so a test can tell which
    decode a sample came from."""
    device = "cpu"

    def decode(self, tok, sample_range=None, **kw):
        import torch
        B, n = tok.shape
        f = n // 7
        full = (torch.arange(2048 * f).unsqueeze(0).repeat(B, 1) + 0 * f).to(torch.int32) + 1_000_000 * f
        return full if sample_range is None else full[:, sample_range[0]:sample_range[1]]


def test_lookahead_policy_bookkeeping(monkeypatch):
    """PIPELINE_REPORT.md:497-505: decode all frames every `frames_per_chunk` new frames, emit only samples with
    >= 5 frames of future context, never twice, flush at the end."""
    import torch
    from tts_inference_b200 import policy
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self, raising=False)
    la = policy.LookaheadStreamingDecoder(_FakeDecoder(), lookahead_frames=5, frames_per_chunk=3)
    got = {"a": [], "b": []}
    total = {"a": 23, "b": 4}
    for step in range(30):
        for s in ("a", "b"):
            if step < total[s]:
                la.push(s, [128266 + step] * 7)
            elif step == total[s]:
                la.finish(s)
        for s, pcm in la.step():
            got[s].append(pcm)
    for s in ("a", "b"):
        cat = np.concatenate(got[s])
        assert cat.size == 2048 * total[s]
        assert (cat % 1_000_000 == np.arange(cat.size)).all()          # every sample exactly once, in order
        frames_of = cat // 1_000_000                                   # which decode emitted it
        t = np.arange(cat.size)
        final = frames_of == total[s]
        assert ((frames_of - 5) * 2048 > t)[~final].all()              # >= 5 frames of context unless it is the flush
    assert policy.stable_samples(4, 5, False) == 0 and policy.stable_samples(4, 5, True) == 8192
    assert not la._streams


@pytest.mark.gpu
def test_lookahead_policy_against_batch_decode(decoder):
    """Streamed audio against the batch decode of the same tokens.  The built-in noise is keyed by (stream, t) and not by
    the decoded length, and 5 frames of lookahead exceed the decoder's receptive field, so the streamed samples are
    BIT-IDENTICAL to the batch decode (the reference, which redraws torch.randn per decode, accepts MSE < 1e-3 and
    correlation > 0.998, PIPELINE_REPORT.md:513-519)."""
    import torch
    from tts_inference_b200 import policy, synth
    lens = [24, 17, 24, 9]                          # streams of different lengths share (and leave) the batched decodes
    tokens = synth.make_tokens(4, 24, seed=9)
    la = policy.LookaheadStreamingDecoder(decoder, lookahead_frames=5, frames_per_chunk=4, seed=3)
    got = [[] for _ in lens]
    for f in range(max(lens) + 1):
        for s, n in enumerate(lens):
            if f < n:
                la.push(s, tokens[s, 7 * f: 7 * f + 7].tolist())
            elif f == n:
                la.finish(s)
        for s, pcm in la.step():
            got[s].append(pcm)
    keys = torch.arange(4, dtype=torch.int32).cuda()
    for s, n in enumerate(lens):
        ref = decoder.decode(torch.from_numpy(np.ascontiguousarray(tokens[s:s + 1, :7 * n])).cuda(), raw_ids=True, seed=3,
                             stream_keys=keys[s:s + 1]).cpu().numpy()[0]
        cat = np.concatenate(got[s])
        assert cat.size == 2048 * n
        mse = np.mean(((cat.astype(np.float64) - ref) / 32768.0) ** 2)
        corr = np.corrcoef(cat.astype(np.float64), ref.astype(np.float64))[0, 1]
        assert mse < 1e-3 and corr > 0.998, (mse, corr)        # the reference's thresholds
        assert np.array_equal(cat, ref), s                       # and in fact bit-identical


@pytest.mark.gpu
def test_stateful_policy_against_batch_decode(decoder):
    """The same streams through the stateful session policy (per-slot, per-stage state in HBM, no re-decode of the
    prefix): bit-identical to the batch decode, streams of different lengths joining and leaving shared steps, slots
    reused by later streams."""
    import torch
    from tts_inference_b200 import policy, synth
    lens = [24, 17, 24, 9, 13, 6]
    starts = [0, 0, 0, 0, 12, 20]                   # the last two streams start later and reuse freed slots
    tokens = synth.make_tokens(len(lens), 24, seed=9)
    sd = policy.StatefulStreamingDecoder(decoder, max_streams=4, window_frames=32, frames_per_chunk=4, seed=3)
    got = [[] for _ in lens]
    for f in range(40):
        for s, n in enumerate(lens):
            g = f - starts[s]
            if 0 <= g < n:
                sd.push(s, tokens[s, 7 * g: 7 * g + 7].tolist())
            elif g == n:
                sd.finish(s)
        for s, pcm in sd.step():
            got[s].append(pcm)
    for s, n in enumerate(lens):
        key = torch.tensor([s], dtype=torch.int32).cuda()
        ref = decoder.decode(torch.from_numpy(np.ascontiguousarray(tokens[s:s + 1, :7 * n])).cuda(), raw_ids=True, seed=3,
                             stream_keys=key).cpu().numpy()[0]
        cat = np.concatenate(got[s])
        assert cat.size == 2048 * n, (s, cat.size)
        assert np.array_equal(cat, ref), s


@pytest.mark.gpu
def test_device_ingest_flush_only_step_and_many_slots():
    """A step that carries no ids at all (n_tok = 0) but finishes streams flushes their remainders; 5000 slots."""
    import torch
    from tts_inference_b200.ingest import DeviceIngest
    S = 5000
    ing = DeviceIngest(S)
    ids = torch.full((S, 16), BASE + 3, dtype=torch.int32).cuda()
    ids[:, 0] = SOS
    wt, ws, tt, ts, tf = ing.step(ids)                                   # 15 ids buffered per stream, nothing ready
    assert wt.shape[0] == 0 and tt.shape[0] == 0
    fin = torch.zeros(S, dtype=torch.uint8).cuda()
    fin[::2] = 1
    wt, ws, tt, ts, tf = ing.step(torch.empty((S, 0), dtype=torch.int32).cuda(), finish=fin)
    assert wt.shape[0] == 0 and tt.shape[0] == S // 2
    assert ts.cpu().tolist() == list(range(0, S, 2)) and (tf == 2).all()
    assert (tt[:, :14] == BASE + 3).all() and (tt[:, 14:] == 0).all()
    st, cnt = ing.state()
    assert (st[::2] == 2).all() and (st[1::2] == 1).all() and (cnt[1::2] == 15).all() and (cnt[::2] == 0).all()
    wt, ws, tt, ts, tf = ing.step(ids[:, 1:14].contiguous())             # 13 more ids: the live streams complete a window
    assert ws.cpu().tolist() == list(range(1, S, 2)) and tt.shape[0] == 0
