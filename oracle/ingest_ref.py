"""TEST INFRASTRUCTURE -- CPU restatement of the steps either side of the decode path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package; the product never does.

* ``ingest_stream`` / ``ingest_steps``: what one stream's LLM token ids turn into, following
  ``generate_audio_tokens`` (vllm_inference/modal_audio_stream.py:313-333 -- skip up to and including the first
  TOKEN_SOS, stop at TOKEN_EOS, yield every other id; the same rule at tensorrt_tts/inference.py:231-241) and
  ``stream_audio`` (:352-396 -- ``snac_code = token_id - TOKEN_AUDIO_BASE``, pop the first 28 codes whenever 28 are
  buffered, at the end emit the remaining whole frames).
  **Pinned**: tests/golden/ingest_golden.json holds what the reference's OWN two functions produce for seeded
  token streams (tests/golden/make_golden_ingest.py runs them out of /root/reference with a stub engine that, like
  vLLM, finishes a request with its stop token -- the reference's own ``break`` at :330 only leaves the inner
  ``for o in out.outputs`` loop and relies on ``stop_token_ids=[TOKEN_EOS]`` (:295) to end the stream; the TensorRT
  variant's plain loop (tensorrt_tts/inference.py:234-241) stops by itself.  Here TOKEN_EOS ends the stream).
* ``last_sos_audio_tokens``: the Hindi/Canopy batch rule (tensorrt_tts/hindi_canopy/inference.py:137-150): everything
  after the LAST TOKEN_SOS, up to TOKEN_EOS.
* ``b64`` / ``wav_bytes``: the stdlib calls the reference's endpoints make (:484; :561-566, :650-657).
"""
from __future__ import annotations

import base64
import io
import wave
from typing import Iterable, List, Sequence, Tuple

TOKEN_SOS = 128257          # modal_audio_stream.py:101
TOKEN_EOS = 128258          # modal_audio_stream.py:102
TOKEN_AUDIO_BASE = 128266   # modal_audio_stream.py:103
FRAME_TOKENS = 7            # modal_audio_stream.py:352
CHUNK_TOKENS = 28           # modal_audio_stream.py:353


def ingest_stream(token_ids: Iterable[int]) -> List[List[int]]:
    """All chunks (lists of ``id - 128266`` codes) ``stream_audio`` hands to ``convert_to_audio`` for one request."""
    chunks: List[List[int]] = []
    buffer: List[int] = []
    found = False
    for t in token_ids:
        if not found:                       # :320-325
            if t == TOKEN_SOS:
                found = True
            continue
        if t == TOKEN_EOS:                  # :328-330
            break
        buffer.append(t - TOKEN_AUDIO_BASE)  # :366-367
        if len(buffer) >= CHUNK_TOKENS:      # :370-372
            chunks.append(buffer[:CHUNK_TOKENS])
            buffer = buffer[CHUNK_TOKENS:]
    rf = len(buffer) // FRAME_TOKENS         # :391-393
    if rf > 0:
        chunks.append(buffer[: rf * FRAME_TOKENS])
    return chunks


class StreamState:
    __slots__ = ("found", "ended", "buffer")

    def __init__(self):
        self.found, self.ended, self.buffer = False, False, []


def ingest_steps(states: Sequence[StreamState], tokens: Sequence[Sequence[int]], finish: Sequence[bool] = None
                 ) -> Tuple[List[Tuple[int, List[int]]], List[Tuple[int, List[int]]]]:
    """The same policy advanced one LLM step for many streams (the shape of the device kernel): ``tokens[s]`` are the
    ids stream s sampled this step, ``finish[s]`` says its generator ended without TOKEN_EOS.  Returns
    (full windows, end-of-stream remainders) as [(stream, raw ids)] in (stream, time) order."""
    full, tails = [], []
    for s, st in enumerate(states):
        if st.ended:
            continue
        fin = bool(finish[s]) if finish is not None else False
        for t in tokens[s]:
            if not st.found:
                if t == TOKEN_SOS:
                    st.found = True
                continue
            if t == TOKEN_EOS:
                fin = True
                break
            st.buffer.append(int(t))
            if len(st.buffer) >= CHUNK_TOKENS:
                full.append((s, st.buffer[:CHUNK_TOKENS]))
                st.buffer = st.buffer[CHUNK_TOKENS:]
        if fin:
            rf = len(st.buffer) // FRAME_TOKENS
            if rf > 0:
                tails.append((s, st.buffer[: rf * FRAME_TOKENS]))
            st.buffer = []
            st.ended = True
    return full, tails


def last_sos_audio_tokens(output_ids: Sequence[int]) -> List[int]:
    """tensorrt_tts/hindi_canopy/inference.py:137-150."""
    sos = [i for i, t in enumerate(output_ids) if t == TOKEN_SOS]
    if not sos:
        return []
    out = []
    for t in output_ids[sos[-1] + 1:]:
        if t == TOKEN_EOS:
            break
        out.append(t)
    return out


def b64(chunk: bytes) -> bytes:
    return base64.b64encode(chunk)          # modal_audio_stream.py:484


def wav_bytes(pcm: bytes, rate: int = 24000) -> bytes:
    buf = io.BytesIO()                      # modal_audio_stream.py:561-566
    with wave.open(buf, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(pcm)
    return buf.getvalue()
