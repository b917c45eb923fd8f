import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_model():
    from oracle import synth_ckpt
    return synth_ckpt.make_model(0)


@pytest.fixture(scope="session")
def state_dict():
    from tts_inference_b200 import synth
    return synth.make_state_dict(0)


@pytest.fixture(scope="session")
def decoder(state_dict):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tts_inference_b200 import SnacDecoder
    return SnacDecoder(state_dict, device=0)
