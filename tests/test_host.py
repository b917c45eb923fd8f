"""CPU tests of the host side: C-ABI surface, loud failure without a GPU, sharding over ranks (gloo)."""
import os
import re
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from tts_inference_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "snacb.h")).read()
    declared = sorted(set(re.findall(r"\b(snacb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libsnacb.so does not export {name}"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.snacb_version() == 100
    assert lib.snacb_samples_out(4, 0) == 8192 and lib.snacb_samples_out(4, _lib.EXTRACT_SLICE) == 2048
    assert lib.snacb_samples_out(2, _lib.EXTRACT_SLICE) == 4096        # not > AUDIO_SLICE_END: all samples kept


def test_no_cpu_fallback(state_dict):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tts_inference_b200 import SnacDecoder, SnacbError, compat
    with pytest.raises(SnacbError, match="no CUDA device"):
        SnacDecoder(state_dict)
    with pytest.raises(RuntimeError):
        compat.convert_to_audio(list(range(28)))       # init_snac() never succeeded


def test_init_snac_never_fabricates_weights(monkeypatch):
    """The reference's init_snac() takes no argument (modal_audio_stream.py:106): the drop-in then needs SNACB_CKPT and
    raises without it -- it must not come back with random weights."""
    from tts_inference_b200 import compat
    monkeypatch.delenv("SNACB_CKPT", raising=False)
    with pytest.raises(RuntimeError, match="no checkpoint"):
        compat.init_snac()
    monkeypatch.setenv("SNACB_CKPT", "/nonexistent/snac_24khz")
    with pytest.raises(FileNotFoundError):
        compat.init_snac()


def test_checkpoint_file_loader_and_fold_cache(tmp_path, state_dict):
    """weights.load_folded: pytorch_model.bin in either weight-norm key style (and the nested directory form) folds to
    the same arrays as the in-memory state dict; the cache entry is keyed by file content and survives corruption."""
    import torch
    from tts_inference_b200 import weights
    want = weights.fold_state_dict(state_dict)
    old = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in state_dict.items()}
    new = {k.replace(".weight_g", ".parametrizations.weight.original0").replace(".weight_v", ".parametrizations.weight.original1"): v
           for k, v in old.items()}
    new["encoder.block.0.bias"] = torch.zeros(4)
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    torch.save(old, tmp_path / "a" / "pytorch_model.bin")
    torch.save(new, tmp_path / "b" / "pytorch_model.bin")
    cache = tmp_path / "cache"
    for src in (tmp_path / "a" / "pytorch_model.bin", tmp_path / "b"):
        got = weights.load_folded(str(src), cache_dir=str(cache))
        assert set(got) == set(want)
        for k in want:
            assert got[k].dtype == np.float32 and got[k].flags["C_CONTIGUOUS"] and np.array_equal(got[k], want[k]), k
    entries = sorted(cache.glob("folded-*.npz"))
    assert len(entries) == 2
    again = weights.load_folded(str(tmp_path / "b"), cache_dir=str(cache))          # served by the cache
    assert all(np.array_equal(again[k], want[k]) for k in want)
    for e in entries:
        e.write_bytes(b"corrupt")                                                   # a damaged entry is rebuilt, not trusted
    again = weights.load_folded(str(tmp_path / "b"), cache_dir=str(cache))
    assert all(np.array_equal(again[k], want[k]) for k in want)
    assert weights.load_folded(str(tmp_path / "a"), cache_dir="")["tail_w"].shape == (1, 64, 7)   # cache disabled


def test_adversarial_synthetic_checkpoints_are_sane():
    """The parity tests' adversarial checkpoints (alpha in {0, +-1e-4, -0.7, 12, 40}, 4x activations) on the oracle:
    finite, tanh not saturated -- otherwise a GPU/oracle comparison on them would prove nothing."""
    import torch
    from oracle import glue_ref, synth_ckpt
    from tts_inference_b200 import synth
    for seed, mode, scale in ((1, "wild", 1.0), (2, "hard", 1.0), (1, "hard", 4.0)):
        sd = synth.make_state_dict(seed, alpha_mode=mode, act_scale=scale)
        al = np.concatenate([v.reshape(-1) for k, v in sd.items() if k.endswith(".alpha")])
        assert (al == 40.0).any() and (al < 0).any() and ((al == 0).any() == (mode == "hard"))
        m = synth_ckpt.make_model(seed, state_dict=sd)
        tok = synth.make_tokens(1, 2, seed=3)
        lv = glue_ref.unpack_np(tok.astype(np.int64) - 128266)
        y = m.decode([torch.from_numpy(x.astype(np.int64)) for x in lv], [torch.from_numpy(n) for n in synth.make_noises(1, 8)])
        assert torch.isfinite(y).all() and 0.05 < float(y.std()) < 0.9 and float((y.abs() > 0.999).float().mean()) < 0.05


def test_product_does_not_import_oracle():
    import subprocess
    import sys
    code = "import sys; import tts_inference_b200, tts_inference_b200.compat, tts_inference_b200.batcher, " \
           "tts_inference_b200.synth, tts_inference_b200.dist; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tts_inference_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_weights_struct_layout(state_dict):
    from tts_inference_b200 import _lib, weights
    folded = weights.fold_state_dict(state_dict)
    w = _lib.make_weights(folded)
    import ctypes as C
    assert C.sizeof(_lib.Weights) == 8 * (9 + 4 + 4 * (4 + 3 * 6) + 3)
    assert np.ctypeslib.as_array(w.block[3].res[2].pw_b, shape=(64,))[5] == folded["b3.r2.pw_b"][5]


def test_shard_range_partitions():
    from tts_inference_b200.dist import shard_range, stream_owner
    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                s, c = shard_range(n, r, world)
                cover += list(range(s, s + c))
            assert cover == list(range(n))
            sizes = [shard_range(n, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert [stream_owner(i, 8) for i in (0, 7, 8, 4095)] == [0, 7, 0, 7]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from tts_inference_b200 import synth
    from tts_inference_b200.dist import max_over_ranks, shard_range, sum_over_ranks
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    tokens = synth.make_tokens(37, 4, seed=5)
    s, c = shard_range(tokens.shape[0], rank, world)
    mine = tokens[s:s + c]
    checksum = float(mine.astype(np.int64).sum())
    dist.barrier()
    t = max_over_ranks(10.0 + rank)             # stands in for the per-rank device time
    total = sum_over_ranks(float(c))
    csum = sum_over_ranks(checksum)
    q.put((rank, c, t, total, csum))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    """N>1 path of bench.py on CPU: shard, barrier, max-over-ranks timing, whole-job count."""
    import torch.multiprocessing as mp
    from tts_inference_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [19, 18]
    assert all(r[2] == 11.0 for r in res)                       # max over ranks
    assert all(r[3] == 37.0 for r in res)                       # every stream decoded exactly once
    assert all(r[4] == float(synth.make_tokens(37, 4, seed=5).astype(np.int64).sum()) for r in res)


@pytest.mark.parametrize("C", [64, 128, 256])
def test_chain_span_schedule(C):
    """Schedule of the fused chain kernel's in-place prologue (host logic, no GPU): emulate the kernel's protocol
    on integers -- every warp first fetches the 3 rows before/after its spans, then all warps rewrite their rows
    in place in arbitrary order -- and check that every row whose result is needed at that layer is produced from
    the layer's ORIGINAL inputs (no read-after-overwrite hazard), for dilations 1, 3, 9."""
    import ctypes as Ct
    from tts_inference_b200 import _lib
    lib = _lib.load()
    buf = (Ct.c_int16 * (3 * 16 * 4 * 3))()
    rc = lib.snacb_debug_chain_spans(C, buf, len(buf))
    rows, nw = rc & 0xFFFF, rc >> 16
    assert rows % 128 == 0 and 256 <= rows <= 1024 and nw in (8, 16)
    sp = np.frombuffer(buf, dtype=np.int16).reshape(3, 16, 4, 3)
    assert (sp[:, nw:, :, 1] == 0).all()
    need_lo = {1: 4, 3: 13, 9: 40}          # first row whose result is consumed downstream, per dilation
    rng = np.random.default_rng(0)
    for l, d in enumerate((1, 3, 9)):
        for kc in range(C // 64):
            x = rng.integers(1, 1 << 30, size=rows).astype(np.int64)       # layer input (one value per row)
            f = lambda r, src: int(sum((j + 2) * (src[r + (j - 3) * d] if 0 <= r + (j - 3) * d < rows else 0)
                                       for j in range(7)))                 # stands in for the 7-tap op
            want = {r: f(r, x) for r in range(rows)}
            work = x.copy()
            spans = [(w, k, *sp[l, w, k]) for w in range(16) for k in range(4) if sp[l, w, k, 1] > 0 and sp[l, w, k, 2] == kc]
            pre = {}
            for (w, k, r0, noct, _) in spans:                               # phase 1: pre-reads
                assert r0 % 8 == 0
                head = [work[r0 - (3 - j) * d] if r0 - (3 - j) * d >= 0 else 0 for j in range(3)]
                tail = [work[r0 + (8 * noct + j) * d] if r0 + (8 * noct + j) * d < rows else 0 for j in range(3)]
                pre[(w, k)] = (head, tail)
            written = set()
            for (w, k, r0, noct, _) in sorted(spans, key=lambda s_: rng.random()):   # phase 2, any warp order
                head, tail = pre[(w, k)]
                n = 8 * noct
                get = lambda i: (head[i + 3] if i < 0 else tail[i - n] if i >= n else
                                 (work[r0 + i * d] if 0 <= r0 + i * d < rows else 0))
                win = [get(i) for i in range(-3, 3)]
                for q in range(noct):
                    raw = [get(8 * q + kk + 3) for kk in range(8)]          # the octet's 8 loads come first
                    for kk in range(8):
                        win = win[-6:] + [raw[kk]]
                        r = r0 + (8 * q + kk) * d
                        if 0 <= r < rows:
                            assert r not in written
                            written.add(r)
                            work[r] = sum((j + 2) * win[j] for j in range(7))
            for r in range(need_lo[d], rows - need_lo[d]):     # the schedule may skip rows nobody consumes
                assert r in written and work[r] == want[r], (C, d, kc, r)
    per_warp = sp[:, :nw, :, 1].sum(axis=2)
    assert per_warp.max() - per_warp.min() <= 1                             # balanced to one octet
