"""GPU parity of the SNAC encode path (csrc/encoder.cu) against oracle/snac_enc_ref.py, and the codes -> tokens -> decode
round trip through the decode path."""
import numpy as np
import pytest
import torch

from oracle import glue_ref
from tts_inference_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def enc_pair():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.snac_enc_ref import SnacEncodeRef
    from tts_inference_b200.encoder import SnacEncoder
    sd = synth.make_encoder_state_dict(0)
    m = SnacEncodeRef().eval()
    m.load_snac_state_dict({k: torch.from_numpy(np.ascontiguousarray(v).copy()) for k, v in sd.items()})
    return SnacEncoder(sd, device=0), m, sd


@pytest.mark.parametrize("B,n", [(1, 2048), (3, 2048 * 5 - 777), (2, 2048 * 16), (4, 100)])
def test_encode_matches_oracle(enc_pair, B, n):
    """Latent within 2e-4 (fp32 both sides); codes identical to the oracle's except where the oracle's two nearest codes
    are closer than 1e-5 (an argmax over fp32 distances cannot be pinned tighter across summation orders) -- and on these
    inputs there is no such tie, so they are identical."""
    enc, m, _ = enc_pair
    audio = synth.make_audio(B, n, seed=3)
    taps = {}
    ref = m.encode(torch.from_numpy(audio)[:, None, :], taps)
    c0, c1, c2, z, dist = enc.encode(torch.from_numpy(audio).cuda(), return_latent=True, return_dist=True)
    zr = taps["z"].permute(0, 2, 1).numpy()
    assert z.shape == zr.shape
    assert float(np.abs(z.cpu().numpy() - zr).max()) <= 2e-4
    total = mism = 0
    for l, (got, want) in enumerate(zip((c0, c1, c2), ref)):
        got = got.cpu().numpy().astype(np.int64)
        want = want.numpy()
        assert got.shape == want.shape and got.min() >= 0 and got.max() < 4096
        d = taps[f"dist{l}"].numpy()                                    # [B, T, 4096]
        bad = np.argwhere(got != want)
        for b, t in bad:                                                # every mismatch must be a near-tie in the oracle
            assert abs(d[b, t, got[b, t]] - d[b, t, want[b, t]]) < 1e-5, (l, b, t)
        best = np.take_along_axis(d, want[..., None], axis=-1)[..., 0]
        assert np.allclose(dist[l].cpu().numpy(), best, atol=2e-5)
        total += want.size; mism += len(bad)
    assert mism <= max(1, total // 200), (mism, total)


def test_tokens_round_trip_through_the_decoder(enc_pair, decoder):
    """encode -> pack_tokens gives the 7-ids-per-frame layout the decode path unpacks (modal_audio_stream.py:156-188):
    unpack(pack(codes)) == codes, the packed ids equal the reference formula, and decoding them runs and is identical to
    decoding the same ids built on the host.  (With random-init weights decode(encode(x)) is not a reconstruction of x --
    that is a property of trained weights, which are not available here.)"""
    enc, m, _ = enc_pair
    audio = torch.from_numpy(synth.make_audio(2, 2048 * 4, seed=8)).cuda()
    c0, c1, c2 = enc.encode(audio)
    tok = enc.pack_tokens(c0, c1, c2, raw_ids=True)
    assert tok.shape == (2, 28)
    u0, u1, u2 = decoder.unpack(tok, raw_ids=True)
    assert torch.equal(u0, c0) and torch.equal(u1, c1) and torch.equal(u2, c2)
    c0h, c1h, c2h = (x.cpu().numpy() for x in (c0, c1, c2))
    want = np.zeros((2, 28), dtype=np.int64)
    for f in range(4):
        fr = [c0h[:, f], c1h[:, 2 * f], c2h[:, 4 * f], c2h[:, 4 * f + 1], c1h[:, 2 * f + 1], c2h[:, 4 * f + 2], c2h[:, 4 * f + 3]]
        for p in range(7):
            want[:, 7 * f + p] = fr[p] + 4096 * p + 128266
    assert np.array_equal(tok.cpu().numpy(), want)
    lv = glue_ref.unpack_np(want - 128266)
    assert all(np.array_equal(a, b) for a, b in zip(lv, (c0h, c1h, c2h)))
    pcm = decoder.decode(tok, raw_ids=True, seed=1)
    assert pcm.shape == (2, 8192)
    assert torch.equal(pcm, decoder.decode(torch.from_numpy(want.astype(np.int32)).cuda(), raw_ids=True, seed=1))
    codes_only = enc.pack_tokens(c0, c1, c2, raw_ids=False)
    assert torch.equal(codes_only + 128266, tok)


def test_encode_empty_and_bad_arguments(enc_pair):
    import ctypes as C
    from tts_inference_b200 import _lib
    enc, _, _ = enc_pair
    lib = _lib.load()
    assert lib.snacb_encode_frames(0) == 0 and lib.snacb_encode_frames(1) == 1 and lib.snacb_encode_frames(2049) == 2
    assert lib.snacb_encode(enc._e, None, 0, 0, 0, None, None, None, None, None, None) == 0          # nothing to do
    assert lib.snacb_encode(enc._e, None, 1, 2048, 2048, None, None, None, None, None, None) == -1   # null pointers
    a = torch.zeros((1, 64), dtype=torch.float32).cuda()
    assert lib.snacb_encode(enc._e, a.data_ptr(), 1, 64, 32, None, None, None, None, None, None) == -1  # stride < n
    assert lib.snacb_encoder_last_error(enc._e)
    l0 = enc.launches()
    c0, c1, c2 = enc.encode(a)                                             # silence: one frame of padding
    assert c0.shape == (1, 1) and enc.launches() - l0 == 1 + 4 * 7 + 1 + 3
