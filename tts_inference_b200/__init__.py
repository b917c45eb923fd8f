"""tts_inference_b200 -- B200-native SNAC 24 kHz decode for Orpheus/Canopy token streams.

One hot path of Demon-Sheriff/tts-inference, rebuilt from scratch as hand-written sm_100a CUDA
behind a C ABI (include/snacb.h): token window -> SNAC codes -> VQ decode -> conv decoder ->
int16 PCM.  ``compat`` mirrors the reference's helper names; ``api.SnacDecoder`` is the batched
device-tensor interface; ``batcher.WindowBatcher`` packs many streams into one launch.
"""
from .api import SnacDecoder, SnacbError, StreamingSession  # noqa: F401

__all__ = ["SnacDecoder", "SnacbError", "StreamingSession"]
