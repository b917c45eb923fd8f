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
    _check_chain_span_schedule(C, 0, False)


@pytest.mark.parametrize("C", [64, 128, 256])
def test_chain_span_schedule_carry_top_tile(C):
    """A tile that is not the first of its strip owns its rows from row 0 on: every dilation class starts at or above
    row 0 and its first span takes the three class rows above row 0 from the carry the previous tile left."""
    _check_chain_span_schedule(C, 0, True)


@pytest.mark.parametrize("C,own_end,carry", [(64, 48, False), (64, 456, False), (64, 240, False), (128, 150, False), (128, 200, False),
                                             (128, 64, False), (256, 184, False), (256, 120, False), (256, 138, False),
                                             (128, 215, False), (256, 41, False), (64, 385, False), (64, 8, True), (64, 440, True),
                                             (128, 208, True), (128, 1, True), (256, 200, True), (256, 77, True), (128, 211, True)])
def test_chain_span_schedule_short_last_tile(C, own_end, carry):
    """The same emulation for the shorter schedule of a row range's LAST tile, which owns rows up to `own_end` only (round 2:
    it used to run as a full tile): every row up to its right halo is still produced from the layer's original inputs."""
    _check_chain_span_schedule(C, own_end, carry)


def _check_chain_span_schedule(C, own_end, carry):
    """Schedule of the fused chain kernel's in-place prologue (host logic, no GPU): emulate the kernel's protocol
    on integers -- every warp first fetches the 3 rows before/after its spans (from the carry for the top span of a class
    of a carry-top tile), then all warps rewrite their rows in place in arbitrary order -- and check that every row whose
    result is needed at that layer is produced from the layer's ORIGINAL inputs (no read-after-overwrite hazard), for
    dilations 1, 3, 9."""
    import ctypes as Ct
    from tts_inference_b200 import _lib
    lib = _lib.load()
    buf = (Ct.c_int16 * (3 * 16 * 4 * 4))()
    rc = lib.snacb_debug_chain_spans_ex(C, own_end, 1 if carry else 0, buf, len(buf))
    assert rc > 0
    rows, nw = rc & 0xFFFF, rc >> 16
    assert rows % 128 == 0 and 256 <= rows <= 1024 and nw in (8, 16)
    sp = np.frombuffer(buf, dtype=np.int16).reshape(3, 16, 4, 4)
    assert (sp[:, nw:, :, 1] == 0).all()
    if not carry:
        assert (sp[..., 3] == 0).all()
    need_lo = {1: 4, 3: 13, 9: 40}          # halo-top tile: first row whose result is consumed downstream, per dilation
    rng = np.random.default_rng(0)
    for l, d in enumerate((1, 3, 9)):
        for kc in range(C // 64):
            x = rng.integers(1, 1 << 30, size=rows).astype(np.int64)       # layer input (one value per row)
            above = rng.integers(1, 1 << 30, size=3 * d).astype(np.int64)  # the 3 d rows above the tile (previous tile)
            src_at = lambda r: (int(x[r]) if 0 <= r < rows else (int(above[r + 3 * d]) if carry and -3 * d <= r < 0 else 0))
            want = {r: sum((j + 2) * src_at(r + (j - 3) * d) for j in range(7)) for r in range(rows)}   # stands in for the 7-tap op
            work = x.copy()
            spans = [(w, k, *[int(v) for v in sp[l, w, k]]) for w in range(16) for k in range(4)
                     if sp[l, w, k, 1] > 0 and sp[l, w, k, 2] == kc]
            pre = {}
            for (w, k, r0, noct, _, top) in spans:                          # phase 1: pre-reads
                assert r0 % 8 == 0
                if top > 0:
                    r_nn = r0 + (top - 1) * d
                    assert 0 <= r_nn < d and r_nn - d < 0                   # the class's first row inside the tile
                    head = [int(above[r_nn - (3 - j) * d + 3 * d]) for j in range(3)]
                else:
                    head = [work[r0 - (3 - j) * d] if r0 - (3 - j) * d >= 0 else
                            (int(above[r0 - (3 - j) * d + 3 * d]) if carry else 0) for j in range(3)]
                tail = [work[r0 + (8 * noct + j) * d] if r0 + (8 * noct + j) * d < rows else 0 for j in range(3)]
                pre[(w, k)] = (head, tail)
            written = set()
            for (w, k, r0, noct, _, top) in sorted(spans, key=lambda s_: rng.random()):   # phase 2, any warp order
                head, tail = pre[(w, k)]
                n = 8 * noct
                k0 = top - 1

                def get(i):
                    if top > 0 and k0 - 3 <= i < k0:
                        return head[i - (k0 - 3)]
                    if i < 0:
                        return head[i + 3] if top == 0 else 0
                    if i >= n:
                        return tail[i - n]
                    return work[r0 + i * d] if 0 <= r0 + i * d < rows else 0
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
            lo = 0 if carry else need_lo[d]
            need_hi = (own_end if own_end else rows - 40) + (36, 27, 0)[l]
            for r in range(lo, need_hi):                       # the schedule may skip rows nobody consumes
                assert r in written and work[r] == want[r], (C, d, kc, r)
            if own_end:                                        # and a short tile does skip: nothing far past its right halo
                assert max(written) < min(rows, need_hi + 8 * d)
    per_warp = sp[:, :nw, :, 1].sum(axis=2)
    assert (per_warp.max(axis=1) - per_warp.min(axis=1) <= 1).all()         # every layer balanced to one octet


def test_chain_strip_plan():
    """Strips of the chain kernel (host logic): the tiles of a plan own exactly the row range, for many range lengths."""
    import ctypes as Ct
    from tts_inference_b200 import _lib
    lib = _lib.load()
    out = (Ct.c_int32 * 4)()
    for C, rows in ((64, 512), (128, 256), (256, 256)):
        own_h, own_c = rows - 80, rows - 40
        for t_n in list(range(1, 1200, 7)) + [1024, 2048, 4096, 8192, 2592, 1296, 16384, 131072]:
            for S, slots in ((1, 148), (64, 148), (1024, 296), (5000, 148)):
                assert lib.snacb_debug_chain_plan(C, t_n, S, slots, out) == 0
                K, sps, lst, lrows = out[0], out[1], out[2], out[3]
                assert 1 <= K <= 24 and sps >= 1 and 1 <= lst <= K
                SR = own_h + (K - 1) * own_c
                full_last = own_h if lst == 1 else own_c
                assert 0 <= lrows < full_last
                owned = (sps - 1) * SR + (own_h + (lst - 2) * own_c if lst > 1 else 0) + (lrows if lrows else full_last)
                assert owned == t_n, (C, t_n, S, K, sps, lst, lrows)
    lib.snacb_debug_chain_plan(256, 1024, 1024, 148, out)
    assert list(out) == [5, 1, 5, 200]              # BASELINE configs[1] at B = 1024, block 1: 5 tiles per stream instead of 6


@pytest.mark.parametrize("policy", [0, 1])
def test_batcher_queue_from_many_producer_threads(policy):
    """The sharded batcher (csrc/batcher.cpp) without a decoder: 8 producer threads push 200 streams token by token /
    in ragged pieces while a consumer takes ready windows concurrently; every stream's windows equal the oracle's
    (stream_audio's 28-code chunks + end-of-stream remainder, modal_audio_stream.py:352-396; or the sliding 28/7
    rule), in time order; a push after end is refused until the id is forgotten."""
    import threading
    from oracle import glue_ref
    from tts_inference_b200 import SnacbError, synth
    from tts_inference_b200.batcher import WindowBatcher
    b = WindowBatcher(None, policy=policy, raw_ids=True, max_windows=4096)
    streams = {1000 + 7 * i: synth.make_tokens(1, 12, seed=i)[0][: 84 - (i % 9)] for i in range(200)}
    ids = list(streams)
    got = {sid: [] for sid in ids}
    done = threading.Event()

    def producer(k):
        rng = np.random.default_rng(k)
        mine = ids[k::8]
        pos = {s: 0 for s in mine}
        live = list(mine)
        while live:
            s = live[int(rng.integers(len(live)))]
            n = int(rng.integers(1, 4)) if k % 2 else 1
            b.push(s, streams[s][pos[s]: pos[s] + n])
            pos[s] += n
            if pos[s] >= len(streams[s]):
                b.end(s)
                live.remove(s)

    def consumer():
        while True:
            fin = done.is_set()
            for sid, tok in b.take(257):
                got[sid].append(tok.tolist())
            if fin and b.pending() == 0:
                return

    th = [threading.Thread(target=producer, args=(k,)) for k in range(8)]
    ct = threading.Thread(target=consumer)
    ct.start()
    for t in th:
        t.start()
    for t in th:
        t.join()
    done.set()
    ct.join()
    for sid, toks in streams.items():
        want = glue_ref.stream_chunks(toks.tolist()) if policy == 0 else glue_ref.sliding_windows(toks.tolist())
        assert got[sid] == want, sid
    with pytest.raises(SnacbError):
        b.push(ids[0], [1, 2, 3])                      # ended: refused, not a silent new stream
    b.forget(ids[0])
    b.push(ids[0], [1, 2, 3])                          # forgotten: the id is free again
    with pytest.raises(SnacbError):
        b.flush()                                      # no decoder attached
    b.close()


@pytest.mark.parametrize("C", [64, 128])
def test_ws_chain_span_schedule(C):
    """Data schedule of the warp-specialised chain kernel's prologue (kernels_chain_ws.cu), emulated on integers in the
    kernel's order: blocks of a layer one after the other; per block every warp first fetches the 3 rows before / after
    its spans (from the block, from the carry copy of the previous block's last 27 rows, or from the untouched next
    block), the warps then copy the block's last rows to the carry buffer, and only then rewrite their rows in place in
    arbitrary order.  Every row of the tile must come out of the layer's ORIGINAL values (no read-after-overwrite)."""
    import ctypes as Ct
    from tts_inference_b200 import _lib
    lib = _lib.load()
    buf = (Ct.c_int16 * (3 * 16 * 4 * 3))()
    rc = lib.snacb_debug_chain_ws_spans(C, buf, len(buf))
    rows, nw = rc & 0xFFFF, rc >> 16
    assert rows % 128 == 0 and nw in (7, 8)
    nb = rows // 128
    sp = np.frombuffer(buf, dtype=np.int16).reshape(3, 16, 4, 3)
    assert (sp[:, nw:, :, 1] == 0).all()
    rng = np.random.default_rng(1)
    for l, d in enumerate((1, 3, 9)):
        quads = sp[l, :nw, :, 1].sum(axis=1)
        assert quads.max() - quads.min() <= 1                               # balanced to one quad (4 steps)
        for kc in range(C // 64):
            x = rng.integers(1, 1 << 30, size=rows + 128).astype(np.int64)  # + rows past the tile (garbage the halo absorbs)
            f = lambda r, src: int(sum((j + 2) * (src[r + (j - 3) * d] if r + (j - 3) * d >= 0 else 0) for j in range(7)))
            want = {r: f(r, x) for r in range(rows)}
            work = x.copy()
            carry = {}
            for b in range(nb):
                base = 128 * b
                spans = [(w, k, int(sp[l, w, k, 0]), int(sp[l, w, k, 1])) for w in range(nw) for k in range(4)
                         if sp[l, w, k, 1] > 0 and sp[l, w, k, 2] == kc]
                pre = {}
                for (w, k, r0, nq) in spans:                                # phase 1: pre-reads
                    head = []
                    for j in range(3):
                        rh = r0 - (3 - j) * d
                        if rh >= 0:
                            head.append(work[base + rh])
                        elif b > 0:
                            assert rh >= -27
                            head.append(carry[(b - 1) & 1][27 + rh])
                        else:
                            head.append(0)
                    tail = [work[base + r0 + (4 * nq + j) * d] for j in range(3)]
                    pre[(w, k)] = (head, tail)
                if b + 1 < nb:
                    carry[b & 1] = work[base + 101: base + 128].copy()
                written = set()
                for (w, k, r0, nq) in sorted(spans, key=lambda s_: rng.random()):   # phase 2 (after the block barrier)
                    head, tail = pre[(w, k)]
                    mem = lambda i: work[base + r0 + i * d]                 # a load from the tile copy, whatever it holds now
                    win = [0] + head + [mem(0), mem(1), mem(2)]
                    noct = (nq + 1) // 2
                    for o in range(noct):
                        full = nq - 2 * o >= 2
                        raw = [mem(8 * o + kk + 3) for kk in range(8)]      # the octet's 8 loads come first
                        if o == noct - 1:
                            if full:
                                raw[5:8] = tail
                            else:
                                raw[1:4] = tail
                        for kk in range(8 if full else 4):
                            win = win[-6:] + [raw[kk]]
                            r = r0 + (8 * o + kk) * d
                            if r < 128:
                                assert base + r not in written
                                written.add(base + r)
                                work[base + r] = sum((j + 2) * win[j] for j in range(7))
                assert written == set(range(base, base + 128)), (C, d, kc, b)
            for r in range(rows - 27):                                      # the last 3d rows see the garbage past the tile
                assert work[r] == want[r], (C, d, kc, r)


def test_ws_chain_protocol():
    """The mbarrier protocol of kernels_chain_ws.cu under random timing: 7 prologue warps, 8 epilogue warps and the IO
    thread as coroutines stepping in random order over several tiles.  mbarriers are modelled with their real semantics
    (a waiter sees only the PARITY of the completed-phase count), so both failure modes of a mis-designed protocol show
    up: a deadlock, or a wait that returns for the wrong phase (checked against the true phase counter).  Data
    dependencies are checked on a shadow state: P(l, b) needs S1_l of blocks b and b+1, MMA needs the operand, E needs the
    MMA, a TMA refill needs the store, the next tile's NoiseBlock MMA needs the refill and a drained accumulator."""
    import random
    NB, NP, NE = 4, 7, 8

    class Bar:
        def __init__(self, count):
            self.count, self.pending, self.done = count, count, 0
        def arrive(self):
            self.pending -= 1
            assert self.pending >= 0
            if self.pending == 0:
                self.pending, self.done = self.count, self.done + 1
        def test(self, parity):                       # mbarrier.try_wait.parity
            return (self.done & 1) != parity

    for seed in range(40):
        rnd = random.Random(seed)
        n_tiles = rnd.choice([1, 2, 3, 5])
        ld = [Bar(1) for _ in range(NB)]; mma = [Bar(1) for _ in range(NB)]; a_ = [Bar(NP) for _ in range(NB)]
        s1 = [Bar(NE) for _ in range(NB)]; out = [Bar(NE) for _ in range(NB)]; tile_bar = Bar(1)
        s_tile = [0, None]
        # shadow state: what each block of the tile copy / TMEM holds
        state = {"copy": [("y", 0)] * NB, "mma_issued": [(-1, -1)] * NB}
        log = {"p": set(), "e": set(), "mma": set(), "store": set(), "load": {(0, b) for b in range(NB)}}
        for b in range(NB):
            ld[b].arrive()                            # first tile's loads land at some point: model as landed

        def wait(bar, parity, want_done):
            while not bar.test(parity):
                yield
            assert bar.done >= want_done, ("wait returned for an earlier phase", bar.done, want_done)
            assert bar.done <= want_done + 1, ("waiter lags two phases: parity would alias", bar.done, want_done)

        def p_warp(w):
            n = 0
            tile = 0
            while tile < n_tiles:
                for l in range(3):
                    for b in range(NB):
                        if b == 0:
                            yield from wait(s1[0], (3 * n + l) & 1, 3 * n + l + 1)
                        if b + 1 < NB:
                            yield from wait(s1[b + 1], (3 * n + l) & 1, 3 * n + l + 1)
                        assert (n, l, b) in log["e"] and (b + 1 == NB or (n, l, b + 1) in log["e"])
                        yield                          # pre-read + carry copy, named barrier, compute
                        log["p"].add((n, l + 1, b, w))
                        a_[b].arrive()
                yield from wait(tile_bar, n & 1, n + 1)
                tile = s_tile[(n + 1) & 1]
                n += 1

        def e_warp(w):
            n = 0
            tile = 0
            while tile < n_tiles:
                for ph in range(4):
                    for b in range(NB):
                        yield from wait(mma[b], ph & 1, 4 * n + ph + 1)
                        assert (n, ph, b) in log["mma"]
                        yield
                        if ph < 3:
                            if w == 0:
                                log["e"].add((n, ph, b))
                            s1[b].arrive()
                        else:
                            if w == 0:
                                log["e"].add((n, 3, b))
                            out[b].arrive()
                yield from wait(tile_bar, n & 1, n + 1)
                tile = s_tile[(n + 1) & 1]
                n += 1

        def io():
            mt = ml = mb = st = sb = 0
            n_local, more, claimed = 1, True, False
            next_free = 1
            while mt < n_local or st < n_local:
                if mt < n_local:
                    if ml == 0:
                        ready = ld[mb].test(mt & 1)
                        if ready:
                            assert ld[mb].done == mt + 1 and (mt, mb) in log["load"]
                            assert mt == 0 or (mt - 1, 3, mb) in log["e"]           # accumulator drained
                    else:
                        ready = a_[mb].test((3 * mt + ml - 1) & 1)
                        if ready:
                            assert a_[mb].done == 3 * mt + ml
                            assert all((mt, ml, mb, w) in log["p"] for w in range(NP))
                    if ready:
                        log["mma"].add((mt, ml, mb))
                        mma[mb].arrive()               # tcgen05.commit, modelled as immediate
                        if ml == 1 and mb == 0 and not claimed:
                            nt = next_free if more else n_tiles
                            next_free += 1
                            more = nt < n_tiles
                            s_tile[(mt + 1) & 1] = nt
                            tile_bar.arrive()
                            if more:
                                n_local += 1
                            claimed = True
                        mb += 1
                        if mb == NB:
                            mb, ml = 0, ml + 1
                            if ml == 4:
                                ml, mt, claimed = 0, mt + 1, False
                yield
                if st < n_local and st <= mt and not (st == mt and ml < 3) and out[sb].test(st & 1):
                    assert out[sb].done == st + 1 and (st, 3, sb) in log["e"]
                    log["store"].add((st, sb))
                    if st + 1 < n_local:
                        log["load"].add((st + 1, sb))
                        ld[sb].arrive()
                    sb += 1
                    if sb == NB:
                        sb, st = 0, st + 1
                yield

        procs = [p_warp(w) for w in range(NP)] + [e_warp(w) for w in range(NE)] + [io()]
        alive = list(range(len(procs)))
        steps = 0
        while alive:
            i = rnd.choice(alive) if rnd.random() < 0.9 else alive[0]
            try:
                next(procs[i])
            except StopIteration:
                alive.remove(i)
            steps += 1
            assert steps < 2_000_000, f"deadlock (seed {seed}, {n_tiles} tiles)"
        assert len(log["store"]) == n_tiles * NB


@pytest.mark.parametrize("chain_mask", [0b1110, 0b1100, 0b0000])
def test_session_frontier_table(chain_mask):
    """snacb_session_step's frontier (rows of every stage that are final once F frames are known) against a brute-force
    dependency walk over the decoder's layers: stem depthwise k7, ConvTranspose1d (k = 2s, stride s, padding s/2: output t
    reads inputs floor((t + s/2) / s) and the one before), NoiseBlock 1x1, ResidualUnits with dilations 1/3/9 (k7),
    tail conv k7.  Frontiers must be monotone in F and everything final with F frames must stay final with more."""
    import ctypes as C
    from tts_inference_b200 import _lib
    lib = _lib.load()
    strides = [8, 8, 4, 2]

    def brute(F):
        L = 4 * F

        def first_invalid(pred, hi):
            t = 0
            while t < hi and pred(t):
                t += 1
            return t
        rows = []
        v = first_invalid(lambda t: t + 3 < L, L + 8)
        rows.append(v)
        for bi, s in enumerate(strides):
            vin = v
            # whole input rows m: both rows an output phase of m may read, m - 1 .. m + 1, must be final
            y = first_invalid(lambda t: (t + s // 2) // s < vin and t // s + 1 < vin, (vin + 2) * s)
            nz = y
            r0 = first_invalid(lambda t: t + 3 < nz, nz + 1)
            r1 = first_invalid(lambda t: t + 9 < r0, r0 + 1)
            r2 = first_invalid(lambda t: t + 27 < r1, r1 + 1)
            if chain_mask >> bi & 1:
                rows += [y, 0, 0, 0, r2]
            else:
                rows += [y, nz, r0, r1, r2]
            v = r2
        rows.append(first_invalid(lambda t: t + 3 < v, v + 1))
        return rows
    prev = None
    for F in list(range(0, 12)) + [31, 64]:
        out = (C.c_int32 * 22)()
        n = lib.snacb_debug_session_frontier(F, chain_mask, out, 22)
        assert n == 22
        got = list(out)
        assert got == brute(F), (F, got, brute(F))
        if prev is not None:
            assert all(a <= b for a, b in zip(prev, got))
        prev = got
    # the emitted samples lag the newest token by the receptive field only (2.5 frames), not by a 5-frame lookahead
    out = (C.c_int32 * 22)()
    lib.snacb_debug_session_frontier(20, chain_mask, out, 22)
    assert 2048 * 17.4 < out[21] < 2048 * 18
