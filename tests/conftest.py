import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def weights():
    from loco_asr_b200.synth import synth_state_dict
    return synth_state_dict(seed=0)


@pytest.fixture(scope="session")
def encoder(weights):
    """One finalized encoder shared by the GPU tests (loads the in-tree libloco_asr.so; no fallback)."""
    import torch
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return LocoSpeechT5Encoder.from_state_dict(weights, device="cuda:0")


@pytest.fixture(scope="session")
def debug_encoder(weights):
    """The LOCO_DEBUG build (libloco_asr_debug.so): the product kernels plus the cross-check kernels and the knobs that select
    them (loco_debug_set).  Only the tests that switch kernels or stop after a layer use it; parity tests run on `encoder`."""
    import torch
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return LocoSpeechT5Encoder.from_state_dict(weights, device="cuda:0", debug=True)
