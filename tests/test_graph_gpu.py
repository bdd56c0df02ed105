"""Whole-step CUDA graph (unispeech_b200/graphed.py): a replayed step must produce the loss and the gradients of the eager step on
the same batch and span mask, for every replay (the mask is data the graph reads, not structure it baked in)."""
import numpy as np
import pytest
import torch

from oracle import wavlm_oracle as O

pytestmark = pytest.mark.gpu


def _build(dev):
    from unispeech_b200.wavlm import WavLM, WavLMConfig
    cfg = O.tiny_config(pre_ln=True, encoder_layers=3)
    m = WavLM(WavLMConfig(vars(cfg)))
    m.load_state_dict(O.deterministic_state_dict(cfg))
    return m.to(dev).train(), cfg


def test_graphed_step_matches_eager(cuda_device):
    from unispeech_b200.graphed import GraphedForwardBackward
    dev = cuda_device
    m, cfg = _build(dev)
    B, L = 2, 16000
    wav, _ = O.deterministic_waveform(B, L, seed=3)
    wav_host = wav.float().pin_memory()
    R = None

    def loss_fn(x):
        nonlocal R
        if R is None:
            R = O.hash_uniform("probe:graph", tuple(x.shape)).to(dev)
        return (x.float() * R).sum()

    g = GraphedForwardBackward(m, loss_fn, B, L, dev).capture()
    np.random.seed(11)
    seen = []
    for step in range(3):
        loss = g.step(wav_host)
        torch.cuda.synchronize()
        mask = g.mask_host.clone()
        seen.append(mask)
        got_loss, got = float(loss), m.grad_buffer().detach().clone()
        # eager step on the same batch and mask
        m.zero_grad_buffer()
        m._engine.prepared_version = None
        x, _ = m.extract_features(wav.to(dev), padding_mask=None, mask=True, mask_indices=mask)
        ref = loss_fn(x)
        ref.backward()
        torch.cuda.synchronize()
        want = m.grad_buffer().detach().clone()
        dev_mask_equal = bool(torch.equal(g.mask_dev.cpu(), mask))
        g.graph.replay()                       # (diagnostics: the same mask replayed again, and the eager step run again)
        torch.cuda.synchronize()
        replay2 = float(g.loss)
        m.zero_grad_buffer()
        x2, _ = m.extract_features(wav.to(dev), padding_mask=None, mask=True, mask_indices=mask)
        eager2 = float(loss_fn(x2))
        assert abs(got_loss - float(ref)) <= 2e-3 * max(1.0, abs(float(ref))), (step, got_loss, replay2, float(ref), eager2, dev_mask_equal,
                                                                                int(mask.sum()), [int(s_.sum()) for s_ in seen])
        scale = want.abs().max().item()
        assert (got - want).abs().max().item() <= 2e-2 * scale, (step, (got - want).abs().max().item(), scale)
    assert not torch.equal(seen[0], seen[1])   # a new span mask every replay


def test_graphed_step_refuses_what_it_cannot_freeze(cuda_device):
    from unispeech_b200.graphed import GraphedForwardBackward
    from unispeech_b200.wavlm import WavLM, WavLMConfig
    cfg = O.tiny_config(pre_ln=True, encoder_layers=2)
    cfg.dropout = 0.1
    m = WavLM(WavLMConfig(vars(cfg))).to(cuda_device).train()
    with pytest.raises(ValueError):
        GraphedForwardBackward(m, lambda x: x.float().sum(), 2, 16000, cuda_device)
    m.eval()
    pad = torch.zeros(2, 16000, dtype=torch.bool)
    pad[1, 12000:] = True
    with pytest.raises(ValueError):
        GraphedForwardBackward(m, lambda x: x.float().sum(), 2, 16000, cuda_device, padding_mask=pad)
