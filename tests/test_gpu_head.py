"""GPU: the classifier head fused into the encoder's last kernel (loco_set_head / loco_set_head_outputs) against the
reference's IntentClassifier.forward (speech_text/intent_classifier.py:24-50) evaluated on each utterance's own
frames -- on the very last_hidden_state the same call returns (isolates the epilogue: max is bit-exact, mean is the
pooled output itself, attention pooling / logits within fp32 summation-order tolerance 2e-5 relative) and on the CPU
oracle's hidden states (end to end: cosine >= 0.999, intent argmax identical)."""
import numpy as np
import pytest
import torch

import helpers as H
from loco_asr_b200._lib import LocoError
from loco_asr_b200.head import IntentHead
from loco_asr_b200.synth import synth_head
from oracle import speecht5_oracle as O

pytestmark = pytest.mark.gpu
LENGTHS = [400, 6400, 20800, 41200, 64000, 100000]        # T = 1, 19, 64, 128, 199, 312


def reference_forward(x, method, q, w, b):
    """IntentClassifier.forward on one unpadded [1, T, 768] sequence, as the reference writes it."""
    x = x[None]
    if method == "average":
        p = torch.mean(x, dim=1, keepdim=True)
    elif method == "max":
        p = torch.max(x, dim=1, keepdim=True).values
    else:
        z = torch.matmul(x, q.T)
        alpha = torch.softmax(z, dim=1)
        p = torch.matmul(alpha.permute(0, 2, 1), x)
    return p[0, 0], torch.nn.functional.linear(p, w, b)[0, 0]


def head_params(seed=5):
    w, b = synth_head(seed)
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(1, 768, generator=g) * 0.08          # a trained-scale query: z spreads over a few units
    return w, b, q


@pytest.mark.parametrize("method", ["average", "max", "attention"])
def test_fused_head_equals_reference_forward_on_the_returned_hidden_states(encoder, method):
    w, b, q = head_params()
    encoder.set_head(IntentHead(w, b, q, method))
    waves = H.make_waves(LENGTHS, seed=23)
    wave = torch.from_numpy(np.concatenate(waves)).cuda()
    pooled, hidden, info = encoder.encode_packed(wave, [len(x) for x in waves], return_hidden=True, with_head=True)
    torch.cuda.synchronize()
    hp, lg = info["head_pooled"].cpu(), info["logits"].cpu()
    hidden, pooled = hidden.cpu(), pooled.cpu()
    assert lg.shape == (len(waves), 101)
    off = 0
    for u, t in enumerate(info["frames"].tolist()):
        ref_p, ref_l = reference_forward(hidden[off:off + t], method, q, w, b)
        if method == "max":
            assert torch.equal(hp[u], ref_p)
        elif method == "average":
            assert torch.equal(hp[u], pooled[u])
            assert torch.allclose(hp[u], ref_p, rtol=0, atol=2e-5 * float(ref_p.abs().max()))
        else:
            assert torch.allclose(hp[u], ref_p, rtol=0, atol=2e-5 * float(ref_p.abs().max())), (u, t)
        assert torch.allclose(lg[u], ref_l, rtol=0, atol=1e-4 * max(1.0, float(ref_l.abs().max()))), (u, t)
        assert int(lg[u].argmax()) == int(ref_l.argmax())
        off += t
    # the head is off again after the call: a plain call (and the host-buffer path, which never sees the head's tensors)
    # returns only the masked mean, identical bits, and leaves the earlier head outputs alone
    keep_hp, keep_lg = info["head_pooled"].clone(), info["logits"].clone()
    again = encoder.encode_packed(wave, [len(x) for x in waves]).cpu()
    assert torch.equal(again, pooled)
    host = encoder.encode_host(wave.cpu().pin_memory(), [len(x) for x in waves])
    assert torch.equal(host, pooled)
    assert torch.equal(info["head_pooled"], keep_hp) and torch.equal(info["logits"], keep_lg)


@pytest.mark.parametrize("method", ["max", "attention"])
def test_fused_head_against_the_cpu_oracle(encoder, weights, method):
    w, b, q = head_params(7)
    encoder.set_head(method=method, q=q, weight=w, bias=b)
    lengths = [9000, 33000, 48000]
    waves = H.make_waves(lengths, seed=31)
    wave = torch.from_numpy(np.concatenate(waves)).cuda()
    _, hp, lg = encoder.encode_packed(wave, lengths, with_head=True)
    hp, lg = hp.cpu(), lg.cpu()
    for u, x in enumerate(waves):
        ref_p, ref_l = reference_forward(O.encode_utterance(weights, torch.from_numpy(x)), method, q, w, b)
        assert H.cosine(hp[u], ref_p) >= 0.999
        assert H.rel_err(hp[u], ref_p) < 3e-2
        assert int(lg[u].argmax()) == int(ref_l.argmax())


def test_pooling_only_head_and_error_paths(encoder):
    waves = H.make_waves([6400, 12000], seed=3)
    wave = torch.from_numpy(np.concatenate(waves)).cuda()
    ns = [len(x) for x in waves]
    encoder.set_head(method="max")                          # no classifier: pooled vector only
    pooled, hidden, info = encoder.encode_packed(wave, ns, return_hidden=True, with_head=True)
    assert info["logits"] is None
    t0 = int(info["frames"][0])
    assert torch.equal(info["head_pooled"][0], hidden[:t0].max(dim=0).values)
    with pytest.raises(LocoError):
        encoder.set_head(method="attention")                # self_attention without q
    with pytest.raises(ValueError):
        encoder.set_head(method="median")
    with pytest.raises(LocoError):
        encoder.set_head(method="average", weight=torch.zeros(101, 768))     # weight without bias
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    from loco_asr_b200.synth import synth_state_dict
    fresh = LocoSpeechT5Encoder.from_state_dict(synth_state_dict(seed=0), device="cuda:0")
    with pytest.raises(LocoError):
        fresh.encode_packed(wave, ns, with_head=True)       # with_head before set_head
