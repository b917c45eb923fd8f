"""Generate the committed golden fixtures.  Run in the BUILD container only:

    python tests/golden/make_golden.py

It executes the reference's OWN glue functions straight out of /root/reference (the
function definitions are pulled out of the files with ``ast`` so that the module-level
``import modal`` / ``fastapi`` side effects are not triggered; no reference source is copied
into this repository -- only the input/output vectors are stored), with the oracle
``SnacDecodeRef`` standing in for the absent pip package ``snac``:

  * ``convert_to_audio``      vllm_inference/modal_audio_stream.py:132-202
  * ``redistribute_codes`` +
    ``decode_snac``           tensorrt_tts/inference.py:54-112
  * ``redistribute_codes``    tensorrt_tts/hindi_canopy/inference.py:47-60

Outputs (committed): ``glue_golden.json`` (integer unpack vectors, every variant) and
``decode_golden.npz`` (tokens, injected noise seeds, the bytes the reference's helper
returned, the fp32 waveform and per-stage statistics of the oracle).
/root/reference does not exist on the GPU box; tests only read the committed outputs.
"""
from __future__ import annotations

import ast
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import glue_ref, synth_ckpt  # noqa: E402


def extract(path, func_names, const_names):
    """Compile selected top-level functions/constants of a reference file into a namespace."""
    src = open(path).read()
    tree = ast.parse(src)
    keep = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in func_names:
            keep.append(node)
        elif isinstance(node, ast.Assign) and any(
                isinstance(t, ast.Name) and t.id in const_names for t in node.targets):
            keep.append(node)
    mod = ast.Module(body=keep, type_ignores=[])
    ns: dict = {}
    exec(compile(mod, path, "exec"), ns)
    return ns


class _InjectedNoiseModel:
    """Wraps the oracle so that the reference helper's ``snac_model.decode(codes)`` call
    (no noise argument) receives the injected noise; records the codes it was handed."""

    def __init__(self, model, noises=None):
        self.model, self.noises, self.seen = model, noises, None

    def decode(self, codes):
        self.seen = [c.clone() for c in codes]
        return self.model.decode(codes, self.noises)


def main():
    torch.manual_seed(0)
    model = synth_ckpt.make_model(0)

    stream = extract(f"{REF}/vllm_inference/modal_audio_stream.py", {"convert_to_audio"},
                     {"AUDIO_SLICE_START", "AUDIO_SLICE_END", "TOKEN_AUDIO_BASE"})
    trt = extract(f"{REF}/tensorrt_tts/inference.py", {"redistribute_codes", "decode_snac"},
                  {"POSITION_OFFSETS", "FRAME_SIZE", "SAMPLE_RATE"})
    canopy = extract(f"{REF}/tensorrt_tts/hindi_canopy/inference.py", {"redistribute_codes"},
                     {"FRAME_SIZE", "TOKEN_BASE"})
    stream["snac_device"] = "cpu"

    # ---------------- integer glue ----------------
    cases = []
    glue_pcm = {}
    good = synth_ckpt.make_codes(1, 6)[0].tolist()
    bad = (synth_ckpt.make_tokens(1, 8, seed=5, bad_frac=0.2)[0].astype(np.int64) - 128266).tolist()
    code_lists = {
        "empty": [], "six": good[:6], "one_frame": good[:7], "ragged_10": good[:10],
        "window_28": good[:28], "ragged_41": good[:41], "bad_56": bad,
        "all_special": [128257 - 128266] * 14, "all_high": [28672 + 5] * 7,
        "zeros": [0] * 14, "max_codes": [4095 + 4096 * (i % 7) for i in range(28)],
    }
    for name, codes in code_lists.items():
        rec = {"name": name, "codes": codes}
        # reference helper (records the codes handed to snac.decode)
        wrap = _InjectedNoiseModel(model, None)
        stream["snac_model"] = wrap
        out = stream["convert_to_audio"](list(codes), False) if len(codes) < 7 else None
        if len(codes) >= 7:
            t0 = 4 * (len(codes) // 7)
            wrap.noises = [torch.from_numpy(n) for n in synth_ckpt.make_noises(1, t0)]
            out = stream["convert_to_audio"](list(codes), False)
            rec["stream_levels"] = [c[0].tolist() for c in wrap.seen]
            rec["stream_pcm_sha256"] = hashlib.sha256(out).hexdigest()
            rec["stream_pcm_len"] = len(out)
            glue_pcm[f"glue_pcm_{name}"] = np.frombuffer(out, dtype=np.int16).copy()   # the bytes themselves (+-1 LSB checks)
        else:
            rec["stream_returns_none"] = out is None
        if len(codes) >= 7:
            rec["trt_levels"] = [list(x) for x in trt["redistribute_codes"](list(codes))]
        if len(codes) >= 7 and len(codes) % 7 == 0:
            # the canopy variant indexes past the end on ragged input; its caller truncates first
            rec["canopy_levels_raw"] = [list(x) for x in canopy["redistribute_codes"](list(codes))]
        cases.append(rec)
    with open(os.path.join(HERE, "glue_golden.json"), "w") as f:
        json.dump({"source": "reference functions executed by tests/golden/make_golden.py",
                   "cases": cases}, f)

    # ---------------- decode (helper end to end, oracle as snac) ----------------
    B, F_ = 3, 4
    tokens = synth_ckpt.make_tokens(B, F_, seed=20241224, bad_frac=0.02)
    noises = synth_ckpt.make_noises(B, 4 * F_, seed=7)
    pcm_full, pcm_slice, pcm_trt, wave = [], [], [], []
    for b in range(B):
        codes = (tokens[b].astype(np.int64) - 128266).tolist()
        nb = [torch.from_numpy(n[b:b + 1]) for n in noises]
        stream["snac_model"] = _InjectedNoiseModel(model, nb)
        pcm_full.append(np.frombuffer(stream["convert_to_audio"](codes, False), dtype=np.int16))
        pcm_slice.append(np.frombuffer(stream["convert_to_audio"](codes, True), dtype=np.int16))
        l0, l1, l2 = trt["redistribute_codes"](codes)
        pcm_trt.append(np.frombuffer(
            trt["decode_snac"](l0, l1, l2, _InjectedNoiseModel(model, nb), "cpu"), dtype=np.int16))
        lv = glue_ref.unpack_np(np.asarray([codes]))
        wave.append(model.decode([torch.from_numpy(x.astype(np.int64)) for x in lv], nb)[0, 0].numpy())
    # a longer ragged utterance through the full-length path (config 3 shape, F=9)
    F2 = 9
    tok2 = synth_ckpt.make_tokens(1, F2, seed=99)
    n2 = synth_ckpt.make_noises(1, 4 * F2, seed=11)
    stream["snac_model"] = _InjectedNoiseModel(model, [torch.from_numpy(n) for n in n2])
    pcm_long = np.frombuffer(
        stream["convert_to_audio"]((tok2[0].astype(np.int64) - 128266).tolist() + [1, 2, 3], False), dtype=np.int16)
    np.savez_compressed(
        os.path.join(HERE, "decode_golden.npz"),
        tokens=tokens, noise_seed=np.int64(7), pcm_full=np.stack(pcm_full), pcm_slice=np.stack(pcm_slice),
        pcm_trt=np.stack(pcm_trt), wave=np.stack(wave).astype(np.float32),
        tokens_long=tok2, noise_seed_long=np.int64(11), pcm_long=pcm_long, ckpt_seed=np.int64(0), **glue_pcm)
    print("wrote glue_golden.json, decode_golden.npz",
          {k: v.shape for k, v in dict(pcm_full=np.stack(pcm_full), pcm_slice=np.stack(pcm_slice),
                                       pcm_long=pcm_long).items()})


if __name__ == "__main__":
    main()
