"""CPU oracle for the SNAC-24 kHz decode hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product (``tts_inference_b200/``) may import this package.  The only
callers allowed are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the checker / reported
baseline, never as the thing shipped.

PARITY UNPINNED at the ``snac.SNAC.decode`` boundary: the arithmetic of the path lives in
the pip package ``snac`` (github hubertsiuzdak/snac, unpinned by the reference:
``.pip_install("snac")`` vllm_inference/modal_audio_stream.py:58, tensorrt_tts/inference.py:26)
and the HF checkpoint ``hubertsiuzdak/snac_24khz``; neither is present in this image or
reachable (no network), and the reference holds no golden (tokens, noise, waveform)
triple.  ``oracle/snac_ref.py`` restates the published module graph of that package.
What IS pinned against the reference's own code:
  * the integer glue (``oracle/glue_ref.py``) -- checked against the reference's own
    Python functions, executed from /root/reference by ``tests/golden/make_golden.py``;
  * Snake1d / ResidualUnit / DecoderBlock structure -- cross-checked against the DAC
    implementation in ``transformers`` (same lineage) in ``tests/test_oracle.py``;
  * weight-norm semantics -- checked against ``torch.nn.utils.weight_norm``.
"""
