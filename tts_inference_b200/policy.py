"""Streaming policies above the decode path, for many streams at once.

``LookaheadStreamingDecoder`` follows the algorithm the reference documents for its recommended engine
(tensorrt_tts/PIPELINE_REPORT.md:475-511; the class itself is not in the reference tree, so this layer is built from
that description and has no reference code to be pinned against):

  1. buffer ALL audio tokens of a stream;
  2. whenever ``frames_per_chunk`` new complete frames have arrived, decode ALL frames from frame 0 (here: hand all
     frames to the decoder but let it compute only the receptive field of the samples about to be emitted,
     ``snacb_decode_range`` -- same bits, a fraction of the work);
  3. emit only samples that have at least ``lookahead_frames`` (5 = 10 240 samples, ~430 ms) of future context,
     tracking ``samples_emitted`` so that nothing is emitted twice;
  4. at end of stream decode once more and emit everything that is left.

Differences by construction: every stream due for a decode in a step is decoded in ONE batched call per distinct length
(the reference decodes one stream per call), and the NoiseBlock noise comes from the counter RNG keyed by
(seed, block, row, t) -- independent of the decoded length -- so a re-decode of a longer prefix draws the SAME noise for
the samples it shares with the shorter one.  Since 5 frames of lookahead exceed the decoder's receptive field, the
streamed audio is then bit-identical to the batch decode of the whole utterance (tests/test_io.py), where the
reference, redrawing torch.randn on every decode (PIPELINE_REPORT.md:481), reports correlation 0.9987.  Each stream's noise is keyed by
its own ``noise_key`` (snacb_decode_keyed), not by its row in whatever batch it happens to be decoded with.
What is still linear in the prefix per step is block 0 and the stem (~10 % of a decode); reusing their state across steps
is DESIGN.md section 8 "next".
"""
from __future__ import annotations

from typing import Dict, Hashable, Iterable, List, Tuple

import numpy as np

FRAME = 7
SAMPLES_PER_FRAME = 2048


def stable_samples(frames: int, lookahead_frames: int, final: bool) -> int:
    """Samples of a ``frames``-frame decode that may be emitted (PIPELINE_REPORT.md:497-505)."""
    return SAMPLES_PER_FRAME * (frames if final else max(0, frames - lookahead_frames))


class _Stream:
    __slots__ = ("ids", "decoded_frames", "emitted", "done", "key", "flushed")

    def __init__(self, key: int):
        self.key = key
        self.ids: List[int] = []
        self.decoded_frames = 0
        self.emitted = 0
        self.done = False
        self.flushed = False


class LookaheadStreamingDecoder:
    def __init__(self, decoder, lookahead_frames: int = 5, frames_per_chunk: int = 4, raw_ids: bool = True,
                 precision: str = "fp16", seed: int = 0):
        if lookahead_frames < 0 or frames_per_chunk < 1:
            raise ValueError("lookahead_frames >= 0 and frames_per_chunk >= 1")
        self._dec = decoder
        self.lookahead_frames, self.frames_per_chunk = int(lookahead_frames), int(frames_per_chunk)
        self.raw_ids, self.precision, self.seed = raw_ids, precision, int(seed)
        self._streams: Dict[Hashable, _Stream] = {}
        self._next_key = 0

    def _get(self, stream: Hashable) -> _Stream:
        st = self._streams.get(stream)
        if st is None:
            st = self._streams[stream] = _Stream(self._next_key)       # noise key: order of first appearance
            self._next_key = (self._next_key + 1) & 0x7FFFFFFF
        return st

    # ------------------------------------------------------------------ producers
    def push(self, stream: Hashable, ids: Iterable[int]) -> None:
        st = self._get(stream)
        if st.done:
            raise ValueError(f"stream {stream!r} already finished")
        st.ids.extend(int(i) for i in ids)

    def finish(self, stream: Hashable) -> None:
        self._get(stream).done = True

    # ------------------------------------------------------------------ decode tick
    def step(self) -> List[Tuple[Hashable, np.ndarray]]:
        """Decode every stream that is due (``frames_per_chunk`` new frames, or finished) -- one batched call per
        distinct length -- and return [(stream, new int16 samples)] for the streams that have something to emit.
        Finished streams are forgotten once flushed."""
        import torch
        due: Dict[Tuple[int, int, int], List[Hashable]] = {}
        for key, st in self._streams.items():
            frames = len(st.ids) // FRAME
            if frames > 0 and (st.done or frames - st.decoded_frames >= self.frames_per_chunk) and \
                    (frames != st.decoded_frames or st.done):
                end = stable_samples(frames, self.lookahead_frames, st.done)
                st.decoded_frames = frames
                if end > st.emitted:
                    due.setdefault((frames, st.emitted, end), []).append(key)
        out: List[Tuple[Hashable, np.ndarray]] = []
        for (frames, lo, hi), keys in sorted(due.items()):
            tok = np.asarray([self._streams[k].ids[: frames * FRAME] for k in keys], dtype=np.int64)
            tok = np.clip(tok, -(2 ** 31), 2 ** 31 - 1).astype(np.int32)
            nkeys = torch.tensor([self._streams[k].key for k in keys], dtype=torch.int32).cuda(self._dec.device)
            # only the new stable samples [lo, hi) are computed (snacb_decode_range): the prefix is re-read, not re-decoded
            pcm = self._dec.decode(torch.from_numpy(tok).cuda(self._dec.device), raw_ids=self.raw_ids,
                                   seed=self.seed, precision=self.precision, stream_keys=nkeys, sample_range=(lo, hi))
            host = pcm.cpu().numpy()
            for row, k in enumerate(keys):
                out.append((k, host[row]))
                self._streams[k].emitted = hi
        for key in [k for k, st in self._streams.items() if st.done]:
            del self._streams[key]
        return out


class StatefulStreamingDecoder:
    """The same policy interface (push / finish / step) on the stateful session (``snacb_session_*``,
    SURVEY.md section 8(f) row 1): every stream owns a slot whose per-stage activations stay in HBM, a step appends the
    new whole frames and gets back exactly the samples that became final -- the prefix is neither re-read nor re-decoded.
    Emission differs from ``LookaheadStreamingDecoder`` only in WHEN samples appear: here as soon as their receptive field
    is inside the known tokens (a lag of 2.5 frames = 5050 samples) instead of after a fixed 5-frame lookahead; the bytes are the same
    (both equal the batch decode of the finished stream, tests/test_io.py).  Streams at different positions share one
    launch sequence per tick (``snacb_session_step_multi``)."""

    def __init__(self, decoder, max_streams: int, window_frames: int = 32, frames_per_chunk: int = 4, raw_ids: bool = True,
                 precision: str = "fp16", seed: int = 0):
        if frames_per_chunk < 1:
            raise ValueError("frames_per_chunk >= 1")
        self._dec = decoder
        self._sess = decoder.open_session(max_streams, window_frames, raw_ids=raw_ids, precision=precision)
        self.frames_per_chunk, self.seed = int(frames_per_chunk), int(seed)
        self._streams: Dict[Hashable, _Stream] = {}
        self._slot: Dict[Hashable, int] = {}
        self._free = list(range(max_streams - 1, -1, -1))
        self._next_key = 0

    @property
    def session(self):
        return self._sess

    def push(self, stream: Hashable, ids: Iterable[int]) -> None:
        st = self._streams.get(stream)
        if st is None:
            if not self._free:
                raise RuntimeError("no free slot: max_streams streams are active")
            st = self._streams[stream] = _Stream(self._next_key)
            self._next_key = (self._next_key + 1) & 0x7FFFFFFF
            self._slot[stream] = self._free.pop()
        if st.done:
            raise ValueError(f"stream {stream!r} already finished")
        st.ids.extend(int(i) for i in ids)

    def finish(self, stream: Hashable) -> None:
        if stream in self._streams:
            self._streams[stream].done = True

    def step(self) -> List[Tuple[Hashable, np.ndarray]]:
        """One decode tick.  Every stream that has ``frames_per_chunk`` new whole frames (or is finished) is served; streams
        past their third frame share ONE launch sequence per distinct number of new frames, wherever they are in their
        utterances and whichever slots they hold (``snacb_session_step_multi``); younger streams are grouped by exact
        position; a finished stream is flushed (``final``) by one ranged decode of its window."""
        import torch
        cap = self._sess.max_frames - 16                       # frames one step may add to a non-empty window
        multi: Dict[Tuple[int, int], List[Hashable]] = {}      # (position class, new frames) -> streams
        finals: Dict[Tuple[int, int], List[Tuple[int, Hashable]]] = {}
        for key, st in self._streams.items():
            avail = len(st.ids) // FRAME                       # whole frames not yet handed to the session
            new = min(avail, cap)
            final = st.done and new == avail
            if final:
                finals.setdefault((st.decoded_frames, new), []).append((self._slot[key], key))
            elif new >= self.frames_per_chunk or (st.done and new > 0):
                multi.setdefault((-1 if st.decoded_frames >= 3 else st.decoded_frames, new), []).append(key)
        out: List[Tuple[Hashable, np.ndarray]] = []

        def tokens_of(keys, new):
            tok = np.asarray([self._streams[k].ids[:new * FRAME] for k in keys], dtype=np.int64)
            tok = np.clip(tok.reshape(len(keys), new * FRAME), -(2 ** 31), 2 ** 31 - 1).astype(np.int32)
            nkeys = torch.tensor([self._streams[k].key for k in keys], dtype=torch.int32).cuda(self._dec.device)
            return torch.from_numpy(tok).cuda(self._dec.device), nkeys

        def account(keys, new, host, final):
            for row, k in enumerate(keys):
                st = self._streams[k]
                st.decoded_frames += new
                del st.ids[:new * FRAME]                       # the session holds what it still needs
                st.emitted += host.shape[1]
                st.flushed = final
                if host.shape[1]:
                    out.append((k, host[row]))

        for (_, new), keys in sorted(multi.items(), key=lambda kv: kv[0]):
            tok, nkeys = tokens_of(keys, new)
            pcm = self._sess.step_multi([self._slot[k] for k in keys], tok, seed=self.seed, stream_keys=nkeys)
            account(keys, new, pcm.cpu().numpy(), False)
        for (have, new), members in sorted(finals.items(), key=lambda kv: kv[0]):
            for slot, k in sorted(members):                    # end of stream: once per stream, one ranged decode each
                tok, nkeys = tokens_of([k], new)
                pcm = self._sess.step(slot, tok, final=True, seed=self.seed, stream_keys=nkeys)
                account([k], new, pcm.cpu().numpy(), True)
        for key in [k for k, st in self._streams.items() if st.done and st.flushed]:
            slot = self._slot.pop(key)
            self._sess.reset(slot, 1)
            self._free.append(slot)
            del self._streams[key]
        return out
