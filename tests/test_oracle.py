"""CPU tests: the oracle against the reference's own vectors and against independent code.

* integer glue  -> tests/golden/glue_golden.json, produced by EXECUTING the reference's own
  functions (tests/golden/make_golden.py);
* helper end-to-end (slice, int16, bytes) -> tests/golden/decode_golden.npz, same origin;
* Snake1d / ResidualUnit / DecoderBlock structure -> the DAC implementation in transformers
  (shared lineage);
* weight-norm fold -> torch._weight_norm.
"""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import glue_ref, snac_ref, synth_ckpt
from tts_inference_b200 import synth, weights

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def glue_cases():
    with open(os.path.join(GOLD, "glue_golden.json")) as f:
        return json.load(f)["cases"]


def test_glue_matches_reference_vectors(glue_cases):
    seen = 0
    for c in glue_cases:
        codes = c["codes"]
        if "stream_returns_none" in c:
            assert c["stream_returns_none"] is True
            assert glue_ref.unpack_stream(codes) is None            # modal_audio_stream.py:153-154
            continue
        got = glue_ref.unpack_stream(codes)
        assert [list(x) for x in got] == c["stream_levels"], c["name"]
        assert [list(x) for x in glue_ref.unpack_trt(codes)] == c["trt_levels"], c["name"]
        if "canopy_levels_raw" in c:          # reference fn does not clamp; its caller clamps a level that is out of range
            clamped = [[min(4095, max(0, v)) for v in lv] for lv in c["canopy_levels_raw"]]
            assert [list(x) for x in glue_ref.unpack_canopy(codes)] == clamped, c["name"]
        # all variants agree after clamping, and equal the vectorised form used by the GPU tests
        n = (len(codes) // 7) * 7
        l0, l1, l2 = glue_ref.unpack_np(np.asarray([codes[:n]], dtype=np.int64))
        assert [l0[0].tolist(), l1[0].tolist(), l2[0].tolist()] == c["stream_levels"] == c["trt_levels"]
        assert c["stream_pcm_len"] == 2 * 2048 * (len(codes) // 7)
        seen += 1
    assert seen >= 8


def test_unpack_np_random_vs_loops():
    tok = synth.make_tokens(5, 6, seed=3, bad_frac=0.1).astype(np.int64) - 128266
    l0, l1, l2 = glue_ref.unpack_np(tok)
    for b in range(tok.shape[0]):
        a = glue_ref.unpack_trt(tok[b].tolist())
        assert (l0[b].tolist(), l1[b].tolist(), l2[b].tolist()) == tuple(a)


def test_pcm16_truncates_like_reference():
    x = torch.tensor([-1.7 / 32767, 1.7 / 32767, 0.99999, -0.99999, 1.0, -1.0, 2.0, -2.0, 0.0])
    t = glue_ref.pcm16_torch(x).numpy()
    n = glue_ref.pcm16_numpy(x.numpy())
    assert t.tolist()[:2] == [-1, 1]
    assert np.array_equal(t[:6], n[:6])           # identical inside [-1, 1] (tanh output range)


def test_stream_buffer_policy():
    codes = list(range(28 * 2 + 15))
    chunks = glue_ref.stream_chunks(codes)
    assert [len(c) for c in chunks] == [28, 28, 14]
    assert chunks[0] == codes[:28] and chunks[2] == codes[56:70]
    assert glue_ref.stream_chunks(list(range(6))) == []
    wins = glue_ref.sliding_windows(list(range(42)))
    assert [w[0] for w in wins] == [0, 7, 14] and all(len(w) == 28 for w in wins)


def test_rng_known_answer():
    # splitmix64 reference vector (seed 0 -> first output)
    from tts_inference_b200.synth import _splitmix64
    assert int(_splitmix64(np.array([0], dtype=np.uint64))[0]) == 0xE220A8397B1DCDAF
    n = synth.rng_normal(1, 2, 200000)
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1.0) < 0.01


def test_fold_matches_torch_weight_norm():
    g = torch.Generator().manual_seed(0)
    for shape in [(6, 4, 3), (4, 6, 16), (8, 1, 7)]:
        v = torch.randn(shape, generator=g)
        gg = torch.rand(shape[0], 1, 1, generator=g) + 0.5
        ref = torch._weight_norm(v, gg, 0).numpy()
        got = weights.fold_weight_norm(gg.numpy(), v.numpy())
        assert np.abs(ref - got).max() < 1e-6


def test_fold_both_key_styles(state_dict):
    a = weights.fold_state_dict(state_dict)
    sd2 = {k.replace("weight_g", "parametrizations.weight.original0").replace("weight_v", "parametrizations.weight.original1"): v
           for k, v in state_dict.items()}
    b = weights.fold_state_dict(sd2)
    assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)
    m = snac_ref.SnacDecodeRef()
    m.load_snac_state_dict({k: torch.from_numpy(v) for k, v in sd2.items()})


def test_convtranspose_weight_norm_is_per_input_channel(oracle_model):
    ct = oracle_model.decoder.model[2].block[1]
    assert ct.weight_g.shape == (1024, 1, 1) and ct.weight_v.shape == (1024, 512, 16)


def test_structure_against_dac():
    dac = pytest.importorskip("transformers.models.dac.modeling_dac")
    torch.manual_seed(0)
    # Snake1d
    s_ref, s_dac = snac_ref.Snake1d(12), dac.Snake1d(12)
    alpha = torch.rand(1, 12, 1) * 2 + 0.3
    s_ref.alpha.data.copy_(alpha); s_dac.alpha.data.copy_(alpha)
    x = torch.randn(2, 12, 50)
    assert torch.equal(s_ref(x), s_dac(x))
    # ResidualUnit (DAC has dense k7 convs: compare with groups=1)
    for dil in (1, 3, 9):
        ru, du = snac_ref.ResidualUnit(12, dil, groups=1), dac.DacResidualUnit(12, dil)
        with torch.no_grad():
            du.snake1.alpha.copy_(ru.block[0].alpha); du.snake2.alpha.copy_(ru.block[2].alpha)
            du.conv1.weight.copy_(ru.block[1].weight); du.conv1.bias.copy_(ru.block[1].bias)
            du.conv2.weight.copy_(ru.block[3].weight); du.conv2.bias.copy_(ru.block[3].bias)
            assert torch.allclose(ru(x), du(x), atol=1e-6)
    # DecoderBlock without NoiseBlock: Snake -> ConvTranspose1d(k=2s, stride s, pad ceil(s/2)) -> 3 units
    cfg = dac.DacConfig(decoder_hidden_size=16, upsampling_ratios=[4])
    db_dac = dac.DacDecoderBlock(cfg, stride=4, stride_index=0)
    db = snac_ref.DecoderBlock(16, 8, 4, noise=False, groups=1)
    with torch.no_grad():
        db_dac.snake1.alpha.copy_(db.block[0].alpha)
        db_dac.conv_t1.weight.copy_(db.block[1].weight); db_dac.conv_t1.bias.copy_(db.block[1].bias)
        for i, du in enumerate((db_dac.res_unit1, db_dac.res_unit2, db_dac.res_unit3)):
            ru = db.block[2 + i]
            du.snake1.alpha.copy_(ru.block[0].alpha); du.snake2.alpha.copy_(ru.block[2].alpha)
            du.conv1.weight.copy_(ru.block[1].weight); du.conv1.bias.copy_(ru.block[1].bias)
            du.conv2.weight.copy_(ru.block[3].weight); du.conv2.bias.copy_(ru.block[3].bias)
        xin = torch.randn(2, 16, 20)
        a, b = db(xin), db_dac(xin)
    assert a.shape == b.shape == (2, 8, 80) and torch.allclose(a, b, atol=1e-5)


def test_decode_shapes_and_param_count(oracle_model):
    n = sum(p.numel() for k, p in oracle_model.state_dict().items() if ".in_proj." not in k)
    # SURVEY.md section 8c counts 13 121 025 effective parameters; weight-norm stores g and v separately
    eff = sum(v.size for v in weights.fold_state_dict({k: v.numpy() for k, v in oracle_model.state_dict().items()}).values())
    assert eff == 13121025, eff
    codes = [torch.randint(0, 4096, (1, 1)), torch.randint(0, 4096, (1, 2)), torch.randint(0, 4096, (1, 4))]
    y = oracle_model.decode(codes)            # the reference's warm-up shapes, modal_audio_stream.py:121-127
    assert y.shape == (1, 1, 2048) and float(y.abs().max()) <= 1.0
    assert n > eff


def test_golden_decode_reproduces(oracle_model):
    z = np.load(os.path.join(GOLD, "decode_golden.npz"))
    tokens = z["tokens"]
    noises = synth.make_noises(tokens.shape[0], 16, seed=int(z["noise_seed"]))
    lv = glue_ref.unpack_np(tokens.astype(np.int64) - 128266)
    y = oracle_model.decode([torch.from_numpy(x.astype(np.int64)) for x in lv], [torch.from_numpy(n) for n in noises])
    wave = y[:, 0].numpy()
    assert np.abs(wave - z["wave"]).max() < 2e-5
    pcm = glue_ref.pcm16_torch(y[:, 0]).numpy()
    assert np.abs(pcm.astype(np.int32) - z["pcm_full"].astype(np.int32)).max() <= 1
    assert np.array_equal(z["pcm_slice"], z["pcm_full"][:, 2048:4096])        # slice semantics of the helper
    assert np.array_equal(z["pcm_trt"], z["pcm_full"])                       # torch vs numpy int16 conversion
    assert z["pcm_long"].shape == (2048 * 9,)                                # ragged tail (3 extra codes) dropped


def test_oracle_noise_injection_and_randomness(oracle_model):
    codes = [torch.zeros((1, 1), dtype=torch.long), torch.zeros((1, 2), dtype=torch.long), torch.zeros((1, 4), dtype=torch.long)]
    nz = [torch.from_numpy(n) for n in synth.make_noises(1, 4, seed=3)]
    a, b = oracle_model.decode(codes, nz), oracle_model.decode(codes, nz)
    assert torch.equal(a, b)
    c = oracle_model.decode(codes)            # un-injected: fresh randn, as the reference (PIPELINE_REPORT.md:481)
    assert not torch.equal(a, c)


def test_convert_to_audio_restatement(oracle_model):
    codes = (synth.make_codes(1, 5)[0]).tolist()
    nz = [torch.from_numpy(n) for n in synth.make_noises(1, 20, seed=1)]
    full = glue_ref.convert_to_audio(oracle_model, codes + [5, 6], False, nz)
    sl = glue_ref.convert_to_audio(oracle_model, codes + [5, 6], True, nz)
    assert len(full) == 2 * 2048 * 5 and len(sl) == 2 * 2048
    assert sl == full[2 * 2048: 2 * 4096]
    assert glue_ref.convert_to_audio(oracle_model, codes[:6]) is None
    l0, l1, l2 = glue_ref.unpack_trt(codes)
    assert glue_ref.decode_snac(oracle_model, l0, l1, l2, nz) == full


# ------------------------------------------------------------------ encode half
def _enc_oracle(seed=0):
    from oracle.snac_enc_ref import SnacEncodeRef
    from tts_inference_b200 import synth
    sd = synth.make_encoder_state_dict(seed)
    m = SnacEncodeRef().eval()
    m.load_snac_state_dict({k: torch.from_numpy(np.ascontiguousarray(v).copy()) for k, v in sd.items()})
    return m, sd


def test_encoder_oracle_structure_and_consistency(oracle_model):
    """The encode restatement: total parameter count of snac_24khz (encoder + quantizer + decoder = 19.8 M),
    output shapes of SNAC.encode incl. preprocess()'s right padding, and z_q of the quantiser's forward pass equal to
    from_codes(codes) of the DECODE oracle (the two halves share codebooks and out_proj)."""
    from oracle.snac_enc_ref import SnacEncodeRef
    from tts_inference_b200 import synth
    m, sd = _enc_oracle()
    n_enc = sum(p.numel() for p in m.encoder.parameters())
    n_q = sum(p.numel() for p in m.quantizers.parameters())
    n_dec = sum(p.numel() for p in oracle_model.decoder.parameters())
    assert (n_enc, n_q, n_dec) == (6690672, 139824, 13012418) and n_enc + n_q + n_dec == 19842914
    audio = torch.from_numpy(synth.make_audio(2, 2048 * 3 - 700))[:, None, :]
    taps = {}
    codes = m.encode(audio, taps)
    assert [tuple(c.shape) for c in codes] == [(2, 3), (2, 6), (2, 12)]
    assert all(int(c.min()) >= 0 and int(c.max()) < 4096 for c in codes)
    assert taps["z"].shape == (2, 768, 12) and 0.3 < float(taps["z"].std()) < 3.0
    with torch.inference_mode():
        z_q = oracle_model.quantizer.from_codes(codes)
    assert torch.allclose(z_q, taps["z_q"], atol=1e-5)
    # the padded tail is zeros: encoding the explicitly padded signal gives the same codes
    padded = SnacEncodeRef.preprocess(audio)
    assert padded.shape[-1] == 2048 * 3 and all(torch.equal(a, b) for a, b in zip(codes, m.encode(padded)))


def test_encoder_fold_matches_weight_norm():
    from tts_inference_b200.weights import fold_encoder_state_dict
    m, sd = _enc_oracle()
    f = fold_encoder_state_dict(sd)
    conv = m.encoder.block[2].block[4]                         # strided conv of the second EncoderBlock: g * v / ||v||
    w = torch._weight_norm(conv.weight_v, conv.weight_g, 0)
    assert np.allclose(f["enc.b1.conv_w"], w.detach().numpy(), atol=1e-6)
    ip = m.quantizers[2].in_proj
    assert np.allclose(f["in_proj_w2"], torch._weight_norm(ip.weight_v, ip.weight_g, 0).detach().numpy(), atol=1e-6)
    assert f["enc.final_w"].shape == (768, 1, 7) and f["enc.b3.conv_w"].shape == (768, 384, 16)
