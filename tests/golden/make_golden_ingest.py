"""Golden vectors for the token ingest.  Run in the BUILD container only:

    python tests/golden/make_golden_ingest.py

Executes the reference's OWN ``generate_audio_tokens`` and ``stream_audio``
(vllm_inference/modal_audio_stream.py:272-409) and its last-SOS extraction
(tensorrt_tts/hindi_canopy/inference.py:137-150) straight out of /root/reference: the two async generators are pulled
out of the file with ``ast`` (no module-level ``import modal`` side effects, nothing copied into this repository),
``engine`` is a stub that replays a seeded token sequence the way vLLM streams it (cumulative ``o.token_ids``, the
request finishing with the first id listed in ``stop_token_ids`` -- the reference passes ``[TOKEN_EOS]``, :295), and
``convert_to_audio`` is a stub that records the code lists it is handed.  Output: ``ingest_golden.json``
(LLM token ids in, the chunks stream_audio decodes out).  /root/reference does not exist on the GPU box; tests only
read the committed JSON.
"""
from __future__ import annotations

import ast
import asyncio
import json
import os
import sys
import time
import types
import uuid

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SOS, EOS, BASE = 128257, 128258, 128266


def extract_async(path, names, consts):
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body
            if (isinstance(n, ast.AsyncFunctionDef) and n.name in names) or
            (isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id in consts for t in n.targets))]
    ns: dict = {}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns


class _Out:
    def __init__(self, ids):
        self.token_ids = ids


class _Req:
    def __init__(self, ids):
        self.outputs = [_Out(ids)]


class StubEngine:
    """Replays `ids` one token per iteration, cumulative token_ids like vLLM's RequestOutput; like vLLM it finishes
    the request with the stop token (which is part of token_ids)."""

    def __init__(self, ids):
        self.ids, self.served = list(ids), 0

    async def generate(self, prompt, sampling_params, request_id=None):
        stop = set(sampling_params.get("stop_token_ids") or [])
        for i in range(len(self.ids)):
            self.served = i + 1
            yield _Req(self.ids[: i + 1])
            if self.ids[i] in stop:
                return

    async def abort(self, request_id):
        return None


def run_reference(ns, ids):
    chunks = []
    ns["engine"] = StubEngine(ids)
    ns["convert_to_audio"] = lambda codes, extract_slice=False: (chunks.append((list(codes), bool(extract_slice))) or b"\0\0")

    async def drive():
        async for _ in ns["stream_audio"]("prompt", "tara"):
            pass
    asyncio.run(drive())
    return chunks, ns["engine"].served


def make_stream(rng, kind):
    n_audio = int(rng.integers(0, 150))
    audio = [int(BASE + 4096 * (p % 7) + rng.integers(0, 4096)) for p in range(n_audio)]
    pre = [int(rng.integers(0, 128000)) for _ in range(int(rng.integers(0, 5)))]
    if kind == "plain":
        return pre + [SOS] + audio + [EOS] + [int(BASE + 5)] * 3
    if kind == "no_eos":
        return pre + [SOS] + audio
    if kind == "no_sos":
        return pre + audio + [EOS]
    if kind == "double_sos":                  # a second SOS inside the speech is passed on as a (negative) code
        k = n_audio // 2
        return pre + [SOS] + audio[:k] + [SOS] + audio[k:] + [EOS]
    if kind == "specials":                    # out-of-range ids inside the speech ("can happen with Hindi model tokens")
        for _ in range(min(6, n_audio)):
            audio[int(rng.integers(0, n_audio))] = int(rng.choice([128259, 128260, 128261, 156938, 156999, 5, BASE - 1]))
        return pre + [SOS] + audio + [EOS]
    if kind == "eos_first":
        return pre + [EOS, SOS] + audio
    raise ValueError(kind)


def main():
    sys.modules.setdefault("vllm", types.ModuleType("vllm"))
    sp = types.ModuleType("vllm.sampling_params")
    sp.SamplingParams = lambda **kw: kw
    inp = types.ModuleType("vllm.inputs")
    inp.TokensPrompt = lambda **kw: kw
    sys.modules["vllm.sampling_params"], sys.modules["vllm.inputs"] = sp, inp

    path = f"{REF}/vllm_inference/modal_audio_stream.py"
    ns = extract_async(path, {"generate_audio_tokens", "stream_audio"},
                       {"TOKEN_SOS", "TOKEN_EOS", "TOKEN_AUDIO_BASE"})
    ns.update(time=time, uuid=uuid, format_prompt=lambda prompt, voice: [1, 2, 3], print=lambda *a, **k: None)

    # the Hindi / Canopy batch rule: the `if sos_indices: ... else: ...` statement of run_inference, by position
    cpath = f"{REF}/tensorrt_tts/hindi_canopy/inference.py"
    ctree = ast.parse(open(cpath).read())
    stmts = []
    for node in ast.walk(ctree):
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "sos_indices" for t in node.targets):
            stmts.append(node)
        if isinstance(node, ast.If) and isinstance(node.test, ast.Name) and node.test.id == "sos_indices":
            stmts.append(node)
    stmts.sort(key=lambda n: n.lineno)
    assert len(stmts) == 2 and 135 <= stmts[0].lineno <= 145, [s.lineno for s in stmts]
    last_sos_code = compile(ast.Module(body=stmts, type_ignores=[]), cpath, "exec")

    rng = np.random.default_rng(20241224)
    cases = []
    kinds = ["plain", "no_eos", "no_sos", "double_sos", "specials", "eos_first"]
    for i in range(36):
        kind = kinds[i % len(kinds)]
        ids = make_stream(rng, kind)
        chunks, served = run_reference(ns, ids)
        assert all(not sl for _, sl in chunks)                # stream_audio calls extract_slice=False (:375, :396)
        env = {"output_ids": list(ids), "SOS_TOKEN": SOS, "EOS_TOKEN": EOS, "print": lambda *a, **k: None}
        exec(last_sos_code, env)
        cases.append({"kind": kind, "ids": ids, "chunks": [c for c, _ in chunks], "tokens_consumed": served,
                      "last_sos_audio_tokens": env["audio_tokens"]})
    out = os.path.join(HERE, "ingest_golden.json")
    with open(out, "w") as f:
        json.dump({"source": "vllm_inference/modal_audio_stream.py:272-409 generate_audio_tokens + stream_audio; "
                             "tensorrt_tts/hindi_canopy/inference.py:137-150 (executed from /root/reference)",
                   "cases": cases}, f)
    print(out, len(cases), "cases,", sum(len(c["chunks"]) for c in cases), "chunks")


if __name__ == "__main__":
    main()
