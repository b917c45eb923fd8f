"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on identical inputs.

Bars (BASELINE.json north_star): unpacked codes bit-exact; fp32 waveform <= 1e-3 max-abs;
16-bit tensor-core path SNR >= 40 dB; int16 within +-1 LSB (fp32 path vs the quantised oracle waveform).

The tensor-core path's default operand type is fp16 (tcgen05 kind::f16, fp32 accumulate; 52 dB).  precision="bf16" runs
every contraction as bf16 tcgen05 MMAs on exactly split operands with fp16 storage (the bf16x3 path, 47.7 dB): plain bf16
storage and operands measure 34-36 dB on the synthetic checkpoint, below the 40 dB bar, and are kept only behind
SNACB_BF16_PLAIN=1 for A/B.  Both precisions are held to the 40 dB bar."""
import os

import numpy as np
import pytest
import torch

from oracle import glue_ref
from tests._util import oracle_decode, pcm_of, snr_db
from tts_inference_b200 import SnacDecoder, SnacbError, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FP32_TOL = 1e-3      # north_star: fp32 waveform within 1e-3 max-abs
TC_SNR_DB = 40.0     # north_star: 16-bit tensor-core path SNR >= 40 dB (met with fp16 operands)
# precision="bf16" is the bf16x3 path (fp16 storage, bf16 tcgen05 MMAs on exactly split operands):
# 47.7 dB measured, held to the same 40 dB bar.  SNACB_BF16_PLAIN=1 selects the round-1 behaviour (bf16 storage, single
# bf16 operands: 34 dB), kept for A/B only.


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------ integer kernel
@pytest.mark.parametrize("B,ntok,raw,bad", [(1, 28, True, 0.0), (7, 28, True, 0.3), (3, 7, False, 0.2),
                                            (5, 41, True, 0.1), (64, 3584, True, 0.01), (2, 6, True, 0.0)])
def test_unpack_bit_exact(decoder, B, ntok, raw, bad):
    F_ = max(1, (ntok + 6) // 7)
    tok = synth.make_tokens(B, F_, seed=B * 100 + ntok, bad_frac=bad)[:, :ntok]
    if not raw:
        tok = (tok.astype(np.int64) - 128266).astype(np.int32)
    c = decoder.unpack(_cuda(tok), raw_ids=raw)
    if ntok < 7:
        assert all(x.numel() == 0 for x in c)
        return
    codes = tok.astype(np.int64) - (128266 if raw else 0)
    ref = glue_ref.unpack_np(codes)
    for got, want in zip(c, ref):
        assert np.array_equal(got.cpu().numpy(), want)
    # and against the reference's scalar loop variants
    for b in range(min(B, 3)):
        l = glue_ref.unpack_trt(codes[b].tolist())
        assert [x[b].cpu().tolist() for x in c] == [list(v) for v in l]


def test_unpack_extreme_ids(decoder):
    tok = np.array([[2 ** 31 - 1, -2 ** 31, 0, 128266, 128266 + 4095, 128266 + 4096, 128265]], dtype=np.int32)
    c = decoder.unpack(_cuda(tok), raw_ids=True)
    ref = glue_ref.unpack_np(tok.astype(np.int64) - 128266)
    for got, want in zip(c, ref):
        assert np.array_equal(got.cpu().numpy(), want)


# ------------------------------------------------------------------------------------ fp32 path
@pytest.mark.parametrize("B,F_", [(1, 4), (3, 4), (2, 1), (2, 5), (1, 9)])
def test_fp32_waveform_parity(decoder, oracle_model, B, F_):
    tokens = synth.make_tokens(B, F_, seed=11 + F_, bad_frac=0.02)
    noises = synth.make_noises(B, 4 * F_, seed=5)
    ref, _ = oracle_decode(oracle_model, tokens, noises)
    pcm, wave = decoder.decode(_cuda(tokens), raw_ids=True, noise=[_cuda(n) for n in noises], precision="fp32",
                               return_wave=True)
    w = wave.cpu().numpy()
    assert w.shape == ref.shape == (B, 2048 * F_)
    err = np.abs(w - ref)
    assert np.isfinite(w).all(), f"{(~np.isfinite(w)).sum()} non-finite samples, first at {np.argwhere(~np.isfinite(w))[:3]}"
    assert err.max() <= FP32_TOL, f"max-abs {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}, {(err > FP32_TOL).sum()} samples over"
    d = np.abs(pcm.cpu().numpy().astype(np.int32) - pcm_of(ref).astype(np.int32))
    assert d.max() <= 1, f"int16 differs by {d.max()} LSB"
    assert np.array_equal(pcm.cpu().numpy(), pcm_of(w)), "quantiser"   # quantiser itself is exact (truncation)


def test_fp32_stage_taps(decoder, oracle_model):
    tokens = synth.make_tokens(2, 4, seed=3)
    noises = synth.make_noises(2, 16, seed=9)
    _, rt = oracle_decode(oracle_model, tokens, noises, want_taps=True)
    decoder.decode(_cuda(tokens), raw_ids=True, noise=[_cuda(n) for n in noises], precision="fp32", keep_taps=True)
    taps = decoder.taps()
    assert set(rt) <= set(taps)
    for k, r in rt.items():
        assert taps[k].shape == r.shape, k
        assert np.abs(taps[k] - r).max() <= 2e-4 * max(1.0, np.abs(r).max()), k


# ------------------------------------------------------------------------------------ bf16 tensor-core path
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("B,F_", [(1, 4), (5, 4), (2, 1), (2, 5), (1, 9), (40, 4)])
def test_tensorcore_snr(decoder, oracle_model, B, F_, prec):
    bar = TC_SNR_DB
    tokens = synth.make_tokens(B, F_, seed=21 + F_, bad_frac=0.02)
    noises = synth.make_noises(B, 4 * F_, seed=6)
    ref, _ = oracle_decode(oracle_model, tokens, noises)
    pcm, wave = decoder.decode(_cuda(tokens), raw_ids=True, noise=[_cuda(n) for n in noises], precision=prec,
                               return_wave=True)
    w = wave.cpu().numpy()
    assert np.isfinite(w).all()
    assert np.array_equal(pcm.cpu().numpy(), pcm_of(w))
    assert snr_db(ref, w) >= bar, snr_db(ref, w)


def test_tensorcore_fp32_stream_is_at_least_as_good(decoder, oracle_model):
    tokens = synth.make_tokens(3, 4, seed=31)
    noises = synth.make_noises(3, 16, seed=6)
    ref, _ = oracle_decode(oracle_model, tokens, noises)
    nz = [_cuda(n) for n in noises]
    _, a = decoder.decode(_cuda(tokens), raw_ids=True, noise=nz, precision="fp16", return_wave=True)
    _, b = decoder.decode(_cuda(tokens), raw_ids=True, noise=nz, precision="fp16", return_wave=True, stream_fp32=True)
    assert snr_db(ref, b.cpu().numpy()) >= snr_db(ref, a.cpu().numpy()) - 0.5 >= TC_SNR_DB - 0.5


def test_tensorcore_stage_taps(decoder, oracle_model):
    tokens = synth.make_tokens(2, 4, seed=4)
    noises = synth.make_noises(2, 16, seed=8)
    _, rt = oracle_decode(oracle_model, tokens, noises, want_taps=True)
    decoder.decode(_cuda(tokens), raw_ids=True, noise=[_cuda(n) for n in noises], precision="fp16", keep_taps=True,
                   unfused=True)
    taps = decoder.taps()
    for k, r in rt.items():
        assert taps[k].shape == r.shape, k
        assert snr_db(r, taps[k]) >= 45.0, (k, snr_db(r, taps[k]))


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("B,F_", [(3, 4), (2, 5), (1, 13), (300, 4)])
def test_noiseblock_tma_epilogue_is_bit_identical(monkeypatch, B, F_, prec):
    """Block 0's NoiseBlock GEMM takes y in and puts x out by TMA through per-warp staging tiles (k_gemm_tc, whole 32-row
    groups; F_ = 5 / 13 leave ragged groups on the per-row path); SNACB_NO_TMA_EPI=1 keeps every row on the per-row path.
    The arithmetic is the same: identical bits in the NoiseBlock output and in the waveform."""
    sd = synth.make_state_dict(0)
    tokens = _cuda(synth.make_tokens(B, F_, seed=51 + F_))
    kw = dict(raw_ids=True, seed=3, precision=prec, keep_taps=True, return_wave=True)
    dec = SnacDecoder(sd, device=0)
    _, wa = dec.decode(tokens, **kw)
    ta = dec.taps()["b0.noise"].copy()
    monkeypatch.setenv("SNACB_NO_TMA_EPI", "1")
    _, wb = dec.decode(tokens, **kw)
    tb = dec.taps()["b0.noise"]
    assert np.isfinite(ta).all() and np.array_equal(ta, tb)
    assert torch.equal(wa, wb)


@pytest.mark.parametrize("B,F_", [(3, 4), (2, 5), (1, 13), (150, 4)])
def test_block0_residual_by_identity_mma_against_v1_kernel(monkeypatch, B, F_):
    """Block 0's ResidualUnits (k_resunit2<512>) add the residual with an identity MMA on the 128B-swizzled x chunk and
    store through a swizzled staging tile by TMA; SNACB_RES_V1=1 selects the first-generation kernel (residual loaded and
    added in the epilogue, per-row stores).  Same operands, a different accumulation order: the three unit outputs agree
    to rounding in every row and channel (a wrong swizzle phase or row shift would show up as whole rows / chunks off)."""
    sd = synth.make_state_dict(0)
    tokens = _cuda(synth.make_tokens(B, F_, seed=91 + F_))
    nz = [_cuda(n) for n in synth.make_noises(B, 4 * F_, seed=4)]
    dec_a = SnacDecoder(sd, device=0)
    monkeypatch.setenv("SNACB_RES_V1", "1")
    dec_b = SnacDecoder(sd, device=0)
    monkeypatch.delenv("SNACB_RES_V1")
    dec_a.decode(tokens, raw_ids=True, noise=nz, precision="fp16", keep_taps=True)
    ta = dec_a.taps()
    dec_b.decode(tokens, raw_ids=True, noise=nz, precision="fp16", keep_taps=True)
    tb = dec_b.taps()
    for k in ("b0.res0", "b0.res1", "b0.res2"):
        a, b = ta[k].astype(np.float64), tb[k].astype(np.float64)
        assert a.shape == b.shape and np.isfinite(a).all(), k
        assert snr_db(b, a) >= 55.0, (k, snr_db(b, a))
        assert np.abs(a - b).max() <= 2.0 ** -6 * max(1.0, np.abs(b).max()), (k, np.abs(a - b).max())


def test_cta_pair_convtranspose_experiment(monkeypatch):
    """k_convt_ph2 (cta_group::2 CTA pairs; SNACB_EXPERIMENTS build only, measured slower than k_convt_ph): block 2's
    ConvTranspose output against k_convt_ph's on the same input -- the same MMAs' worth of fp32 sums in the same order."""
    from tts_inference_b200 import _lib
    if not hasattr(_lib.load(), "snacb_debug_chain_ws_spans") or os.environ.get("SNACB_EXPERIMENTS") != "1":
        pytest.skip("k_convt_ph2 is an experiment: built only with SNACB_EXPERIMENTS=1")
    sd = synth.make_state_dict(0)
    tokens = _cuda(synth.make_tokens(37, 5, seed=13))
    dec = SnacDecoder(sd, device=0)
    dec.decode(tokens, raw_ids=True, seed=2, keep_taps=True)
    a = dec.taps()["b2.convt"].copy()
    monkeypatch.setenv("SNACB_CONVT_2CTA", "1")
    dec.decode(tokens, raw_ids=True, seed=2, keep_taps=True)
    b = dec.taps()["b2.convt"]
    assert np.isfinite(a).all() and np.array_equal(a, b)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("B,F_", [(3, 4), (2, 5), (1, 13)])
def test_resident_weight_convtranspose_against_generic_gemm(monkeypatch, B, F_, prec):
    """Blocks 2 / 3 run their ConvTranspose1d through k_convt_ph (TMA-store epilogue; ragged last row group by per-row
    stores) / k_convt_res; SNACB_NO_CONVT_RES=1 sends them through the generic 2-tap GEMM kernel.  Same operands, fp32
    accumulation in a different order: the ConvTranspose outputs agree to rounding, every row of every phase."""
    sd = synth.make_state_dict(0)
    tokens = _cuda(synth.make_tokens(B, F_, seed=77 + F_))
    nz = [_cuda(n) for n in synth.make_noises(B, 4 * F_, seed=5)]
    dec_a = SnacDecoder(sd, device=0)
    monkeypatch.setenv("SNACB_NO_CONVT_RES", "1")
    dec_b = SnacDecoder(sd, device=0)
    monkeypatch.delenv("SNACB_NO_CONVT_RES")
    dec_a.decode(tokens, raw_ids=True, noise=nz, precision=prec, keep_taps=True)
    ta = dec_a.taps()
    dec_b.decode(tokens, raw_ids=True, noise=nz, precision=prec, keep_taps=True)
    tb = dec_b.taps()
    for k in ("b2.convt", "b3.convt"):
        a, b = ta[k].astype(np.float64), tb[k].astype(np.float64)
        assert a.shape == b.shape and np.isfinite(a).all(), k
        assert snr_db(b, a) >= (60.0 if prec == "fp16" else 45.0), (k, snr_db(b, a))
        if k == "b2.convt":        # same input to both kernels here (block 3's differs by block 2's rounding noise)
            assert np.abs(a - b).max() <= 2.0 ** (-9 if prec == "fp16" else -6) * max(1.0, np.abs(b).max()), k


@pytest.mark.parametrize("B,F_", [(2, 4), (3, 1), (1, 7)])
def test_fused_chain_block_outputs(decoder, oracle_model, B, F_):
    """The fused NoiseBlock + ResidualUnit chain (blocks 2 and 3): block outputs against the oracle's, and the
    waveform against the per-layer kernels on the same inputs."""
    tokens = synth.make_tokens(B, F_, seed=40 + F_, bad_frac=0.02)
    noises = synth.make_noises(B, 4 * F_, seed=8)
    nz = [_cuda(n) for n in noises]
    _, rt = oracle_decode(oracle_model, tokens, noises, want_taps=True)
    _, wf = decoder.decode(_cuda(tokens), raw_ids=True, noise=nz, precision="fp16", keep_taps=True, return_wave=True)
    taps = decoder.taps()
    for k in ("b0.res2", "b1.res2", "b2.res2", "b3.res2"):
        assert taps[k].shape == rt[k].shape, k
        assert snr_db(rt[k], taps[k]) >= 45.0, (k, snr_db(rt[k], taps[k]))
    _, wu = decoder.decode(_cuda(tokens), raw_ids=True, noise=nz, precision="fp16", unfused=True, return_wave=True)
    assert snr_db(wu.cpu().numpy(), wf.cpu().numpy()) >= 45.0


# ------------------------------------------------------------------------------------ warp-specialised chain kernel
@pytest.fixture(scope="module")
def ws_decoder(state_dict):
    """SNACB_CHAIN_WS=1: blocks 2 and 3 (C = 128 / 64) run kernels_chain_ws.cu (prologue / epilogue / IO warps pipelined
    over the 128-row blocks of a tile) instead of the lock-step k_chain."""
    import os
    from tts_inference_b200 import _lib
    if not _lib.load().snacb_experiments_built():
        pytest.skip("k_chain_ws is an experiment (measured slower): built only with SNACB_EXPERIMENTS=1")
    os.environ["SNACB_CHAIN_WS"] = "1"
    try:
        d = SnacDecoder(state_dict, device=0)          # the switch is read when the handle is created
    finally:
        del os.environ["SNACB_CHAIN_WS"]
    yield d
    d.close()


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("B,F_,sliced", [(1, 4, False), (3, 4, True), (37, 4, False), (2, 16, False), (300, 4, True), (5, 1, False),
                                         (3, 33, False)])
def test_ws_chain_is_bit_identical(decoder, ws_decoder, B, F_, sliced, prec):
    """Every element goes through the same arithmetic in the same order as in k_chain: PCM and waveform are equal bit for
    bit, injected and in-kernel noise, whole windows, the trimmed sliced call, long utterances, many tiles per CTA."""
    tokens = _cuda(synth.make_tokens(B, F_, seed=60 + F_, bad_frac=0.01))
    nz = [_cuda(n) for n in synth.make_noises(B, 4 * F_, seed=9)]
    for kw in (dict(noise=nz), dict(seed=5)):
        p0, w0 = decoder.decode(tokens, raw_ids=True, precision=prec, extract_slice=sliced, return_wave=True, **kw)
        p1, w1 = ws_decoder.decode(tokens, raw_ids=True, precision=prec, extract_slice=sliced, return_wave=True, **kw)
        torch.cuda.synchronize()
        assert torch.equal(p0, p1) and torch.equal(w0, w1)


def test_ws_chain_ranged_and_general_variant(ws_decoder, decoder, monkeypatch):
    tokens = _cuda(synth.make_tokens(4, 9, seed=3))
    full = ws_decoder.decode(tokens, raw_ids=True, seed=6, return_wave=True)
    part = ws_decoder.decode(tokens, raw_ids=True, seed=6, return_wave=True, sample_range=(5000, 11000))
    assert torch.equal(part[0], full[0][:, 5000:11000]) and torch.equal(part[1], full[1][:, 5000:11000])
    assert torch.equal(full[0], decoder.decode(tokens, raw_ids=True, seed=6))
    # the general (fp32 Snake) variant of both kernels on a checkpoint that forbids the alpha fold
    sd = synth.make_state_dict(2, alpha_mode="hard")
    a = SnacDecoder(sd, device=0)
    monkeypatch.setenv("SNACB_CHAIN_WS", "1")
    b = SnacDecoder(sd, device=0)
    monkeypatch.delenv("SNACB_CHAIN_WS")
    assert a.chain_modes() == [0, 1, 1, 1]
    tok = _cuda(synth.make_tokens(6, 4, seed=8))
    ra, rb = a.decode(tok, raw_ids=True, seed=2, return_wave=True), b.decode(tok, raw_ids=True, seed=2, return_wave=True)
    assert torch.equal(ra[0], rb[0]) and torch.equal(ra[1], rb[1])
    a.close(); b.close()


# ------------------------------------------------------------------------------------ adversarial checkpoints
ADVERSARIAL = [(1, "wild", 1.0), (2, "hard", 1.0), (3, "hard", 4.0), (4, "wild", 4.0), (5, "benign", 4.0)]


@pytest.fixture(scope="module", params=ADVERSARIAL, ids=lambda p: f"seed{p[0]}-{p[1]}-x{p[2]:g}")
def adversarial(request):
    """Decoder + oracle on a synthetic checkpoint whose Snake alphas include {0, +-1e-4, -0.7, 12, 40} ("hard": forces the
    general chain variant; "wild": the alpha-folded one stays legal) and / or 4x larger activations -- trained alphas are
    not the benign U(0.3, 3) of the default synthetic checkpoint."""
    from oracle import synth_ckpt
    seed, mode, scale = request.param
    sd = synth.make_state_dict(seed, alpha_mode=mode, act_scale=scale)
    dec = SnacDecoder(sd, device=0)
    yield dec, synth_ckpt.make_model(seed, state_dict=sd), mode
    dec.close()


def test_adversarial_checkpoint_parity(adversarial):
    dec, model, mode = adversarial
    # which chain formulation the handle picked for blocks 1-3 (block 0 runs per-layer kernels)
    assert dec.chain_modes() == ([0, 1, 1, 1] if mode == "hard" else [0, 2, 2, 2])
    for B, F_ in ((2, 4), (1, 7)):
        tokens = synth.make_tokens(B, F_, seed=50 + F_, bad_frac=0.02)
        noises = synth.make_noises(B, 4 * F_, seed=12)
        nz = [_cuda(n) for n in noises]
        ref, _ = oracle_decode(model, tokens, noises)
        pcm, wave = dec.decode(_cuda(tokens), raw_ids=True, noise=nz, precision="fp32", return_wave=True)
        w = wave.cpu().numpy()
        assert np.isfinite(w).all()
        assert np.abs(w - ref).max() <= FP32_TOL, np.abs(w - ref).max()
        assert np.abs(pcm.cpu().numpy().astype(np.int32) - pcm_of(ref).astype(np.int32)).max() <= 1
        for kw in (dict(), dict(unfused=True)):
            _, wh = dec.decode(_cuda(tokens), raw_ids=True, noise=nz, precision="fp16", return_wave=True, **kw)
            wh = wh.cpu().numpy()
            assert np.isfinite(wh).all()
            assert snr_db(ref, wh) >= TC_SNR_DB, (kw, snr_db(ref, wh))
        _, wb = dec.decode(_cuda(tokens), raw_ids=True, noise=nz, precision="bf16", return_wave=True)
        assert np.isfinite(wb.cpu().numpy()).all() and snr_db(ref, wb.cpu().numpy()) >= TC_SNR_DB


def test_adversarial_checkpoint_bit_identities(adversarial):
    """Trimmed / ranged / regrouped decodes stay bit-identical to the full decode on the adversarial checkpoints too."""
    dec, _, _ = adversarial
    tokens = _cuda(synth.make_tokens(5, 6, seed=8, bad_frac=0.01))
    for prec in ("fp16", "bf16"):
        full = dec.decode(tokens, raw_ids=True, seed=3, precision=prec, return_wave=True)
        sl = dec.decode(tokens, raw_ids=True, seed=3, precision=prec, extract_slice=True, return_wave=True)
        assert torch.equal(sl[0], full[0][:, 2048:4096]) and torch.equal(sl[1], full[1][:, 2048:4096])
        part = dec.decode(tokens, raw_ids=True, seed=3, precision=prec, sample_range=(5000, 9000), return_wave=True)
        assert torch.equal(part[0], full[0][:, 5000:9000]) and torch.equal(part[1], full[1][:, 5000:9000])
        one = dec.decode(tokens[2:3].contiguous(), raw_ids=True, seed=3, precision=prec,
                         stream_keys=_cuda(np.array([2], dtype=np.int32)))
        assert torch.equal(one[0], full[0][2])


def test_general_chain_variant_on_the_benign_checkpoint(state_dict, oracle_model, monkeypatch):
    """SNACB_NO_FOLD=1 runs the general (fp32 Snake) chain variant where the folded one is legal: same checkpoint, both
    variants meet the bar and agree with each other to fp16 rounding."""
    monkeypatch.setenv("SNACB_NO_FOLD", "1")
    gen = SnacDecoder(state_dict, device=0)
    monkeypatch.delenv("SNACB_NO_FOLD")
    fold = SnacDecoder(state_dict, device=0)
    assert gen.chain_modes() == [0, 1, 1, 1] and fold.chain_modes() == [0, 2, 2, 2]
    tokens = synth.make_tokens(3, 4, seed=91)
    noises = synth.make_noises(3, 16, seed=2)
    nz = [_cuda(n) for n in noises]
    ref, _ = oracle_decode(oracle_model, tokens, noises)
    _, a = gen.decode(_cuda(tokens), raw_ids=True, noise=nz, return_wave=True)
    _, b = fold.decode(_cuda(tokens), raw_ids=True, noise=nz, return_wave=True)
    assert snr_db(ref, a.cpu().numpy()) >= TC_SNR_DB and snr_db(ref, b.cpu().numpy()) >= TC_SNR_DB
    assert snr_db(a.cpu().numpy(), b.cpu().numpy()) >= 45.0
    gen.close(); fold.close()


@pytest.mark.parametrize("cache", [None, "3"])
def test_tensor_map_cache_eviction_keeps_results_identical(state_dict, monkeypatch, cache):
    """More distinct (batch, frames, range) shapes than the activation tensor-map cache holds (4096 entries, ~20 per
    shape; and a 3-entry cache, which evicts inside every launch sequence): maps are handed out by value, so an eviction
    between two lookups of one launch cannot leave it with a dangling map.  Every decode must
    equal the reference decode of the same rows."""
    if cache:
        monkeypatch.setenv("SNACB_TMAP_CACHE", cache)
    dec = SnacDecoder(state_dict, device=0)
    monkeypatch.delenv("SNACB_TMAP_CACHE", raising=False)
    keys = _cuda(np.arange(64, dtype=np.int32))
    n = 0
    for F_ in ((1, 2, 3) if cache is None else (3,)):
        tokens = _cuda(synth.make_tokens(64, F_, seed=17 + F_))
        rng = (100, 2048 * F_ - 500)
        want = dec.decode(tokens, raw_ids=True, seed=2, stream_keys=keys)
        want_r = dec.decode(tokens, raw_ids=True, seed=2, stream_keys=keys, sample_range=rng)
        for B in range(1, 65, 1 if cache is None else 7):
            for ranged in (False, True):
                got = dec.decode(tokens[:B].contiguous(), raw_ids=True, seed=2, stream_keys=keys[:B].contiguous(),
                                 sample_range=rng if ranged else None)
                assert torch.equal(got, (want_r if ranged else want)[:B]), (F_, B, ranged)
                n += 1
    assert n == (384 if cache is None else 20)
    dec.close()


# ------------------------------------------------------------------------------------ helper semantics
@pytest.mark.parametrize("prec", ["fp32", "fp16", "bf16"])
def test_slice_semantics(decoder, prec):
    tokens = synth.make_tokens(3, 4, seed=1)
    full = decoder.decode(_cuda(tokens), raw_ids=True, seed=3, precision=prec)
    sl = decoder.decode(_cuda(tokens), raw_ids=True, seed=3, precision=prec, extract_slice=True)
    assert full.shape == (3, 8192) and sl.shape == (3, 2048)
    assert torch.equal(sl, full[:, 2048:4096])                    # modal_audio_stream.py:195-196
    two = synth.make_tokens(2, 2, seed=2)                          # 4096 samples: not > AUDIO_SLICE_END -> all kept
    a = decoder.decode(_cuda(two), raw_ids=True, seed=3, precision=prec, extract_slice=True)
    b = decoder.decode(_cuda(two), raw_ids=True, seed=3, precision=prec, extract_slice=False)
    assert a.shape == (2, 4096) and torch.equal(a, b)


@pytest.mark.parametrize("F_,unfused", [(3, False), (6, False), (9, False), (6, True), (17, False)])
def test_sliced_call_trims_dead_samples_bit_identically(decoder, F_, unfused):
    """extract_slice=True computes only the receptive field of samples [2048, 4096) (dead-sample trimming, DESIGN.md
    section 4.2): every utterance length and both kernel paths give exactly the slice of the full decode."""
    tokens = synth.make_tokens(5, F_, seed=100 + F_, bad_frac=0.01)
    noises = [_cuda(n) for n in synth.make_noises(5, 4 * F_, seed=3)]
    full = decoder.decode(_cuda(tokens), raw_ids=True, noise=noises, unfused=unfused)
    sl = decoder.decode(_cuda(tokens), raw_ids=True, noise=noises, extract_slice=True, unfused=unfused)
    assert full.shape == (5, 2048 * F_) and sl.shape == (5, 2048)
    assert torch.equal(sl, full[:, 2048:4096])
    rng = decoder.decode(_cuda(tokens), raw_ids=True, seed=77, extract_slice=True, unfused=unfused)   # in-kernel noise
    rng_full = decoder.decode(_cuda(tokens), raw_ids=True, seed=77, unfused=unfused)
    assert torch.equal(rng, rng_full[:, 2048:4096])


@pytest.mark.parametrize("prec", ["fp16", "bf16", "fp32"])
@pytest.mark.parametrize("B,F_,lo,hi", [(3, 4, 2048, 4096), (2, 4, 0, 100), (2, 4, 8000, 8192), (5, 12, 6144, 14336),
                                        (1, 40, 61440, 71680), (4, 9, 0, 18432), (2, 1, 5, 2043), (37, 6, 4096, 8192),
                                        (2, 64, 100000, 104096), (3, 30, 0, 4096), (3, 30, 57344, 61440), (2, 100, 190000, 190001)])
def test_ranged_decode_equals_slice_of_full_decode(decoder, prec, B, F_, lo, hi):
    """snacb_decode_range: samples [lo, hi) only, their receptive field only -- bit-identical to the full decode's slice."""
    tokens = _cuda(synth.make_tokens(B, F_, seed=40 + F_))
    keys = _cuda(np.arange(100, 100 + B, dtype=np.int32))
    full = decoder.decode(tokens, raw_ids=True, seed=6, precision=prec, stream_keys=keys, return_wave=True)
    part = decoder.decode(tokens, raw_ids=True, seed=6, precision=prec, stream_keys=keys, return_wave=True,
                          sample_range=(lo, hi))
    assert part[0].shape == (B, hi - lo)
    assert torch.equal(part[0], full[0][:, lo:hi]) and torch.equal(part[1], full[1][:, lo:hi])
    with pytest.raises(ValueError):
        decoder.decode(tokens, raw_ids=True, sample_range=(0, 2048 * F_ + 1))


def test_ragged_tail_is_dropped(decoder):
    tokens = synth.make_tokens(2, 5, seed=8)
    a = decoder.decode(_cuda(tokens[:, :31]), raw_ids=True, seed=1)     # 4 frames + 3 stray tokens
    b = decoder.decode(_cuda(np.ascontiguousarray(tokens[:, :28])), raw_ids=True, seed=1)
    assert a.shape == (2, 8192) and torch.equal(a, b)


def test_builtin_noise_matches_counter_rng(decoder, oracle_model):
    """noise=None uses the in-kernel counter RNG keyed by (seed, block, stream, t): same values as synth.make_noises_rng."""
    tokens = synth.make_tokens(3, 4, seed=12)
    ref, _ = oracle_decode(oracle_model, tokens, synth.make_noises_rng(3, 16, seed=1234))
    _, wave = decoder.decode(_cuda(tokens), raw_ids=True, seed=1234, precision="fp32", return_wave=True)
    assert np.abs(wave.cpu().numpy() - ref).max() <= FP32_TOL
    _, w2 = decoder.decode(_cuda(tokens), raw_ids=True, seed=1235, precision="fp32", return_wave=True)
    assert not torch.equal(wave, w2)


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
def test_builtin_noise_is_independent_of_decoded_length(decoder, prec):
    """A longer decode of the same stream redraws the same noise for the shared time steps: samples whose receptive
    field lies inside the shorter prefix come out bit-identical (what streaming policies that re-decode prefixes need)."""
    tokens = synth.make_tokens(3, 12, seed=14)
    long_ = decoder.decode(_cuda(tokens), raw_ids=True, seed=77, precision=prec)
    short = decoder.decode(_cuda(np.ascontiguousarray(tokens[:, :7 * 8])), raw_ids=True, seed=77, precision=prec)
    n = 2048 * (8 - 3)                                   # 3 frames of margin >> receptive field
    assert torch.equal(long_[:, :n], short[:, :n])
    assert not torch.equal(long_[:, 2048 * 7: 2048 * 8], short[:, 2048 * 7:])


@pytest.mark.parametrize("prec", ["fp32", "fp16", "bf16"])
def test_grouping_and_batch_independence(decoder, prec):
    """Streams are independent: a stream's PCM does not depend on batch size or group split."""
    B = 37
    tokens = synth.make_tokens(B, 4, seed=77)
    noises = [_cuda(n) for n in synth.make_noises(B, 16, seed=2)]
    decoder.set_group_bytes(48 << 20)
    big = decoder.decode(_cuda(tokens), raw_ids=True, noise=noises, precision=prec)
    decoder.set_group_bytes(5 * 131072 * 4 * 2)                     # forces groups of a few streams
    small = decoder.decode(_cuda(tokens), raw_ids=True, noise=noises, precision=prec)
    decoder.set_group_bytes(48 << 20)
    assert torch.equal(big, small)
    one = decoder.decode(_cuda(tokens[5:6]), raw_ids=True, noise=[n[5:6].contiguous() for n in noises], precision=prec)
    assert torch.equal(one[0], big[5])


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("B,F_,sliced", [(40, 4, False), (301, 4, True), (3, 64, False), (5, 37, False), (1, 1, False)])
def test_tail_kernels_agree(decoder, monkeypatch, prec, B, F_, sliced):
    """The three tail kernels on the same block-3 output.  k_tail_tc (default for 16-bit activations: channel contraction
    as a tcgen05 GEMM against the hi / lo split of the fp32 weights, diagonal sum over the taps) against the FFMA kernels:
    the same sum in another order -> the waveform agrees to fp32 rounding, the PCM to 1 LSB.  The two FFMA kernels
    (SNACB_TAIL_V1=1: one load per lane and row; =2: persistent, cp.async.bulk ring) share their arithmetic: bit-identical."""
    tokens = _cuda(synth.make_tokens(B, F_, seed=11))
    kw = dict(raw_ids=True, seed=4, precision=prec, extract_slice=sliced, return_wave=True)
    new = decoder.decode(tokens, **kw)
    monkeypatch.setenv("SNACB_TAIL_V1", "1")
    old = decoder.decode(tokens, **kw)
    monkeypatch.setenv("SNACB_TAIL_V1", "2")
    bulk = decoder.decode(tokens, **kw)
    torch.cuda.synchronize()
    assert torch.equal(bulk[0], old[0]) and torch.equal(bulk[1], old[1])
    assert int((new[0] != 0).sum()) > new[0].numel() // 2
    assert float((new[1] - old[1]).abs().max()) <= 4e-6
    assert int((new[0].int() - old[0].int()).abs().max()) <= 1
    assert np.array_equal(new[0].cpu().numpy(), pcm_of(new[1].cpu().numpy()))


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("B,F_,rng", [(24, 4, None), (3, 37, None), (7, 4, "slice"), (5, 23, (9000, 30011)), (4, 9, (100, 2300)),
                                      (2, 64, None), (9, 1, None), (3, 5, (0, 17)), (300, 4, None), (1, 200, None)])
def test_chain_tile_geometries_are_bit_identical(decoder, monkeypatch, prec, B, F_, rng):
    """The chain kernel cuts a stream's row range into strips that one CTA walks in order: the first tile of a strip
    recomputes 40 rows of context above its rows (halo-top), every further tile takes the three class rows above each
    dilation class from its predecessor (carry-top, no halo above), and a range that is not a whole number of tiles ends in
    a SHORT tile (own schedule, blocks / epilogue pieces past its right halo skipped): kernels_chain.cu.  Against the same
    decode with one halo-top tile per strip (SNACB_NO_CARRY=1, the round-1 geometry) and with full last tiles."""
    tokens = _cuda(synth.make_tokens(B, F_, seed=17))
    kw = dict(raw_ids=True, seed=6, precision=prec, return_wave=True)
    if rng == "slice":
        kw["extract_slice"] = True
    elif rng is not None:
        kw["sample_range"] = rng
    new = decoder.decode(tokens, **kw)
    monkeypatch.setenv("SNACB_NO_CARRY", "1")
    mid = decoder.decode(tokens, **kw)
    monkeypatch.setenv("SNACB_NO_SHORT_TILE", "1")
    old = decoder.decode(tokens, **kw)
    monkeypatch.delenv("SNACB_NO_CARRY")
    alt = decoder.decode(tokens, **kw)                     # carry-top tiles, full last tile
    torch.cuda.synchronize()
    for other in (mid, old, alt):
        assert torch.equal(new[0], other[0]) and torch.equal(new[1], other[1])
    assert int((new[0] != 0).sum()) > new[0].numel() // 2


def test_golden_vectors(decoder):
    """Committed fixtures: bytes the REFERENCE's convert_to_audio returned with the oracle as SNAC."""
    z = np.load(os.path.join(GOLD, "decode_golden.npz"))
    tokens = z["tokens"]
    noises = [_cuda(n) for n in synth.make_noises(tokens.shape[0], 16, seed=int(z["noise_seed"]))]
    pcm, wave = decoder.decode(_cuda(tokens), raw_ids=True, noise=noises, precision="fp32", return_wave=True)
    assert np.abs(wave.cpu().numpy() - z["wave"]).max() <= FP32_TOL
    assert np.abs(pcm.cpu().numpy().astype(np.int32) - z["pcm_full"].astype(np.int32)).max() <= 1
    sl = decoder.decode(_cuda(tokens), raw_ids=True, noise=noises, precision="fp32", extract_slice=True)
    assert np.abs(sl.cpu().numpy().astype(np.int32) - z["pcm_slice"].astype(np.int32)).max() <= 1
    tl = z["tokens_long"]
    nl = [_cuda(n) for n in synth.make_noises(1, 36, seed=int(z["noise_seed_long"]))]
    pl = decoder.decode(_cuda(tl), raw_ids=True, noise=nl, precision="fp32")
    assert np.abs(pl.cpu().numpy()[0].astype(np.int32) - z["pcm_long"].astype(np.int32)).max() <= 1
    pb, wb = decoder.decode(_cuda(tokens), raw_ids=True, noise=noises, precision="fp16", return_wave=True)
    assert snr_db(z["wave"], wb.cpu().numpy()) >= TC_SNR_DB


def test_decode_host_equals_device(decoder):
    tokens = synth.make_tokens(9, 4, seed=5)
    a = decoder.decode_host(tokens, raw_ids=True, extract_slice=True, seed=4)
    b = decoder.decode(_cuda(tokens), raw_ids=True, extract_slice=True, seed=4)
    assert np.array_equal(a, b.cpu().numpy())


def test_pipelined_host_boundary(decoder):
    """snacb_decode_host_submit / _wait: same bytes as the blocking call, two steps in flight at most."""
    steps = [synth.make_tokens(33, 4, seed=20 + i) for i in range(5)]
    want = [decoder.decode_host(t, raw_ids=True, seed=9 + i) for i, t in enumerate(steps)]
    toks = [torch.from_numpy(t).pin_memory() for t in steps]
    outs = [torch.zeros((33, 8192), dtype=torch.int16).pin_memory() for _ in steps]
    with pytest.raises(SnacbError):
        decoder.wait_host()                                           # nothing outstanding
    for i in range(len(steps)):
        decoder.submit_host_ptr(toks[i].data_ptr(), 33, 28, outs[i].data_ptr(), raw_ids=True, seed=9 + i)
        if i:
            decoder.wait_host()
            assert np.array_equal(outs[i - 1].numpy(), want[i - 1])
    decoder.submit_host_ptr(toks[0].data_ptr(), 33, 28, outs[0].data_ptr(), raw_ids=True, seed=9)
    with pytest.raises(SnacbError):                                   # a third submit needs a wait first
        decoder.submit_host_ptr(toks[1].data_ptr(), 33, 28, outs[1].data_ptr(), raw_ids=True, seed=10)
    decoder.wait_host()
    decoder.wait_host()
    assert np.array_equal(outs[-1].numpy(), want[-1]) and np.array_equal(outs[0].numpy(), want[0])


def test_decode_host_large_batch_equals_device(decoder):
    """Host-buffer entry point on a batch that spans many tiles per kernel: same bytes as the device-buffer call."""
    tokens = synth.make_tokens(520, 4, seed=15)
    a = decoder.decode_host(tokens, raw_ids=True, seed=6)
    b = decoder.decode(_cuda(tokens), raw_ids=True, seed=6)
    assert a.shape == (520, 8192) and np.array_equal(a, b.cpu().numpy())


def test_empty_and_bad_arguments(decoder):
    from tts_inference_b200 import SnacbError
    e = decoder.decode(torch.empty((0, 28), dtype=torch.int32, device="cuda"), raw_ids=True)
    assert e.shape == (0, 8192)
    assert decoder.decode_host(np.zeros((2, 6), dtype=np.int32)).shape == (2, 0)
    with pytest.raises(SnacbError):
        decoder._check(decoder._lib.snacb_decode(decoder._h, None, 1, 28, 4, 0, None, 0, None, None, None), "null")
    tok = _cuda(synth.make_tokens(2, 4, seed=1))
    out = torch.empty((2, 8192), dtype=torch.int16, device="cuda")
    lib = decoder._lib
    for lo, hi in ((-1, 10), (10, 10), (20, 10), (0, 8193)):        # bad sample ranges are refused, nothing is launched
        assert lib.snacb_decode_range(decoder._h, tok.data_ptr(), 2, 28, 4, 1, None, 0, None, lo, hi, out.data_ptr(), None, None) == -1
    assert lib.snacb_decode_range(decoder._h, tok.data_ptr(), 2, 28, 4, 1, None, 0, None, 0, 8192, out.data_ptr(), None, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(out, decoder.decode(tok, raw_ids=True, seed=0))
    assert lib.snacb_decode_host_submit(decoder._h, None, 2, 28, 4, 1, 0, None) == -1


# ------------------------------------------------------------------------------------ full-size properties
def test_full_size_batch_properties(decoder, oracle_model):
    """BASELINE configs[1]/target size (B=1024 windows): determinism, and sampled windows equal their
    stand-alone decode and meet the SNR bar against the oracle."""
    B = 1024
    tokens = synth.make_tokens(B, 4, seed=20241224)
    tok = _cuda(tokens)
    a = decoder.decode_windows(tok, raw_ids=True, seed=9)
    b = decoder.decode_windows(tok, raw_ids=True, seed=9)
    assert a.shape == (B, 2048) and torch.equal(a, b)
    assert int((a != 0).sum()) > B * 1024
    idx = [0, 511, 1023]
    nz = synth.make_noises_rng(B, 16, seed=9)
    sub = [np.ascontiguousarray(n[idx]) for n in nz]
    ref, _ = oracle_decode(oracle_model, tokens[idx], sub)
    _, w = decoder.decode(_cuda(tokens[idx]), raw_ids=True, noise=[_cuda(n) for n in sub], return_wave=True)
    assert snr_db(ref, w.cpu().numpy()) >= TC_SNR_DB
    _, wfull = decoder.decode(tok, raw_ids=True, seed=9, return_wave=True, extract_slice=True)
    assert snr_db(ref[:, 2048:4096], wfull.cpu().numpy()[idx]) >= TC_SNR_DB - 1.0


def test_chain_schedule_survives_jitter(decoder, state_dict):
    """Race detector for k_chain's in-place prologue (compute-sanitizer racecheck is closed on this pool): with
    SNACB_CHAIN_JITTER=seed every warp waits a pseudo-random 0..4095 cycles after the barrier that follows the neighbour
    pre-reads, so the order in which warps rewrite their rows of the tile copy changes from seed to seed.  A read of
    another warp's rows after that barrier would then see rewritten data in some runs: the output must not change."""
    import os
    tokens = _cuda(synth.make_tokens(96, 4, seed=71))
    long_tok = _cuda(synth.make_tokens(3, 40, seed=72))
    ref = decoder.decode(tokens, raw_ids=True, seed=3, return_wave=True)
    ref_long = decoder.decode(long_tok, raw_ids=True, seed=3)
    for seed in (1, 2, 3, 4):
        os.environ["SNACB_CHAIN_JITTER"] = str(seed)
        try:
            d = SnacDecoder(state_dict, device=0)
        finally:
            del os.environ["SNACB_CHAIN_JITTER"]
        got = d.decode(tokens, raw_ids=True, seed=3, return_wave=True)
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]), seed
        assert torch.equal(d.decode(long_tok, raw_ids=True, seed=3), ref_long), seed
        gb = d.decode(tokens, raw_ids=True, seed=3, precision="bf16")
        assert torch.equal(gb, decoder.decode(tokens, raw_ids=True, seed=3, precision="bf16")), seed
        d.close()
    # negative control: with bit 31 of the seed the kernel re-reads its neighbour rows AFTER the barrier and the delay --
    # the hazard itself.  The detector must notice (the output changes), otherwise the passes above would prove nothing.
    os.environ["SNACB_CHAIN_JITTER"] = str((1 << 31) | 5)
    try:
        d = SnacDecoder(state_dict, device=0)
    finally:
        del os.environ["SNACB_CHAIN_JITTER"]
    bad = d.decode(tokens, raw_ids=True, seed=3, return_wave=True)
    assert not torch.equal(bad[1], ref[1])
    d.close()
