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


@pytest.mark.parametrize("C", [64, 128])
def test_chain_span_schedule_halo_exchange(C):
    """The same emulation for the halo-exchange variant (SNACB_XCH=1): the tile owns every row, the 32 rows above and
    below it hold the neighbouring tiles' boundary rows (only the nearest 3*d matter), nothing outside is readable.
    Every row of the tile must come out of the layer's ORIGINAL inputs, neighbours included."""
    import ctypes as Ct
    from tts_inference_b200 import _lib
    lib = _lib.load()
    buf = (Ct.c_int16 * (3 * 16 * 4 * 3))()
    rc = lib.snacb_debug_chain_spans_x(C, buf, len(buf))
    rows, nw = rc & 0xFFFF, rc >> 16
    assert rows == 512 and nw in (8, 16)
    sp = np.frombuffer(buf, dtype=np.int16).reshape(3, 16, 4, 3)
    PAD = 32
    rng = np.random.default_rng(1)
    for l, d in enumerate((1, 3, 9)):
        for kc in range(C // 64):
            ext = rng.integers(1, 1 << 30, size=rows + 2 * PAD).astype(np.int64)     # virtual rows -PAD .. rows+PAD-1
            ext[:PAD - 3 * d] = -7777; ext[rows + PAD + 3 * d:] = -7777                # stale pad rows: must never matter
            at = lambda r, src: int(src[r + PAD]) if -PAD <= r < rows + PAD else 0
            want = {r: sum((j + 2) * at(r + (j - 3) * d, ext) for j in range(7)) for r in range(rows)}
            work = ext.copy()
            spans = [(w, k, *sp[l, w, k]) for w in range(16) for k in range(4) if sp[l, w, k, 1] > 0 and sp[l, w, k, 2] == kc]
            pre = {}
            for (w, k, r0, noct, _) in spans:                                          # phase 1: pre-reads (guard: >= -PAD)
                assert r0 % 8 == 0 and r0 + 27 >= -PAD                                 # in-loop look-ahead stays inside the pad
                head = [at(r0 - (3 - j) * d, work) if r0 - (3 - j) * d >= -PAD else 0 for j in range(3)]
                tail = [int(work[min(r0 + (8 * noct + j) * d + PAD, rows + 2 * PAD - 1)]) for j in range(3)]   # unguarded over-read
                pre[(w, k)] = (head, tail)
            written = set()
            for (w, k, r0, noct, _) in sorted(spans, key=lambda s_: rng.random()):
                head, tail = pre[(w, k)]
                n = 8 * noct

                def get(i):
                    if i < 0:
                        return head[i + 3]
                    if i >= n:
                        return tail[i - n]
                    r = r0 + i * d
                    if i < 3:                                                           # window init: guarded reads
                        return at(r, work) if r >= -PAD else 0
                    return int(work[min(r + PAD, rows + 2 * PAD - 1)])                  # main loop: unguarded
                win = [get(i) for i in range(-3, 3)]
                for q in range(noct):
                    raw = [get(8 * q + kk + 3) for kk in range(8)]
                    for kk in range(8):
                        win = win[-6:] + [raw[kk]]
                        r = r0 + (8 * q + kk) * d
                        if 0 <= r < rows:
                            assert r not in written
                            written.add(r)
                            work[r + PAD] = sum((j + 2) * win[j] for j in range(7))
            for r in range(rows):
                assert r in written and int(work[r + PAD]) == want[r], (C, d, kc, r)
    per_warp = sp[:, :nw, :, 1].sum(axis=2)
    assert per_warp.max() - per_warp.min() <= 1


@pytest.mark.parametrize("C", [64, 128])
def test_chain2_span_schedule(C):
    """Schedule of the two-group chain kernel (kernels_chain2.cu), emulated on integers in the kernel's order: group 0
    stashes its last 27 rows, pre-reads its in-group neighbours, rewrites its half in place and reads the tails that
    lie below the group boundary LATE (group 1 has not touched its rows yet); then group 1 pre-reads its in-group
    neighbours, takes the heads above the boundary from the stash and rewrites its half.  Every row whose result is
    consumed downstream must come out of the layer's ORIGINAL inputs, for dilations 1, 3, 9."""
    import ctypes as Ct
    from tts_inference_b200 import _lib
    lib = _lib.load()
    buf = (Ct.c_int16 * (3 * 16 * 4 * 4))()
    rows = lib.snacb_debug_chain2_spans(C, buf, len(buf))
    assert rows % 256 == 0 and 256 <= rows <= 1024
    half = rows // 2
    sp = np.frombuffer(buf, dtype=np.int16).reshape(3, 16, 4, 4)
    need_lo = {1: 4, 3: 13, 9: 40}
    rng = np.random.default_rng(1)
    LATE, STASH = 1, 2
    for l, d in enumerate((1, 3, 9)):
        for kc in range(C // 64):
            x = rng.integers(1, 1 << 30, size=rows + 128).astype(np.int64)   # reads may run past the tile (garbage)
            f = lambda r: int(sum((j + 2) * (x[r + (j - 3) * d] if 0 <= r + (j - 3) * d < rows else 0) for j in range(7)))
            work = x.copy()
            written = set()
            stash = {r: work[r] for r in range(half - 27, half)}
            for g in (0, 1):
                spans = [(w, k, *map(int, sp[l, w, k])) for w in range(8 * g, 8 * g + 8) for k in range(4)
                         if sp[l, w, k, 1] > 0 and sp[l, w, k, 2] == kc]
                pre = {}
                for (w, k, r0, noct, _, fl) in spans:                       # pre-reads before the group barrier
                    assert r0 + 3 * d >= -8                                 # 1 KB of slack above the tile
                    hrows = [r0 - (3 - j) * d for j in range(3)]
                    trows = [r0 + (8 * noct + j) * d for j in range(3)]
                    if g == 0:
                        assert not fl & STASH
                        assert r0 + (8 * noct - 1) * d < half               # group 0 never writes below the boundary
                        assert bool(fl & LATE) == (trows[0] >= half)
                    else:
                        assert not fl & LATE and r0 >= half
                        assert bool(fl & STASH) == (hrows[2] < half)
                        if fl & STASH:
                            assert all(half - 27 <= r < half for r in hrows)
                    head = [stash[r] if fl & STASH else (work[r] if r >= 0 else 0) for r in hrows]
                    tail = None if fl & LATE else [work[r] for r in trows]
                    if not fl & STASH:
                        assert all(r not in written for r in hrows if r >= 0)
                    if not fl & LATE:
                        assert all(r not in written for r in trows)
                    pre[(w, k)] = (head, tail)
                for (w, k, r0, noct, _, fl) in sorted(spans, key=lambda s_: rng.random()):   # any warp order
                    head, tail = pre[(w, k)]
                    n = 8 * noct
                    if fl & LATE:                                           # read at the last octet, other group's rows
                        trows = [r0 + (n + j) * d for j in range(3)]
                        assert all(half <= r and r not in written for r in trows)
                        tail = [work[r] for r in trows]
                    get = lambda i: (head[i + 3] if i < 0 else tail[i - n] if i >= n else
                                     (work[r0 + i * d] if 0 <= r0 + i * d else 0))
                    win = [get(i) for i in range(-3, 3)]
                    for q in range(noct):
                        raw = [get(8 * q + kk + 3) if 8 * q + kk + 3 < n + 3 else 0 for kk in range(8)]
                        for kk in range(8):
                            win = win[-6:] + [raw[kk]]
                            r = r0 + (8 * q + kk) * d
                            if 0 <= r < rows:
                                assert r not in written and (r < half) == (g == 0)
                                written.add(r)
                                work[r] = sum((j + 2) * win[j] for j in range(7))
            for r in range(need_lo[d], rows - need_lo[d]):
                assert r in written and work[r] == f(r), (C, d, kc, r)
        for g in (0, 1):
            per_warp = sp[l, 8 * g:8 * g + 8, :, 1].sum(axis=1)
            assert per_warp.max() - per_warp.min() <= 1
