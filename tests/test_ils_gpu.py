"""ILS-HuBERT (unispeech_b200/ils_hubert.py): intermediate-layer masked-prediction heads against the CPU oracle -- the oracle's
hidden states at the predicted layers (encoder with a layer LIST, src/fairseq/models/hubert/ils_hubert.py:167-171), its
`compute_nce` / criterion restatements applied per layer in the reference's layer-major order (ils_hubert.py:213-272,
hubert_criterion.py:52-110), torch autograd for the gradients."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import wavlm_oracle as O

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("pre_ln,separate", [(False, True), (True, True), (True, False)])
def test_ils_hubert_heads_match_oracle(cuda_device, pre_ln, separate):
    from unispeech_b200.ils_hubert import ILSHubertConfig, ILSHubertModel
    dev = cuda_device
    base = O.tiny_config(pre_ln=pre_ln, encoder_layers=3, relative_position_embedding=False, gru_rel_pos=False)
    layers, ncls, fd = [1, 3], [40, 24], 64
    cfg = ILSHubertConfig(dict(vars(base), final_dim=fd, predict_layers=str(layers), separate_label_embeds=separate,
                               untie_final_proj=True, logit_temp=0.1, feature_grad_mult=1.0))
    torch.manual_seed(3)
    m = ILSHubertModel(cfg, ncls)
    sd0 = O.deterministic_state_dict(base)
    missing = m.load_state_dict(sd0, strict=False)
    assert not missing.unexpected_keys
    m = m.to(dev).train()
    sd = {k: v.detach().cpu().double().float().requires_grad_(True) for k, v in m.state_dict(keep_vars=True).items()}
    B, L = 2, 16000
    wav, pmask = O.deterministic_waveform(B, L, seed=2, lengths=[16000, 12000])
    T = O.num_frames(L, base)
    np.random.seed(4)
    fpm = O.frame_padding_mask(pmask, T)
    mask = m.apply_mask(B, T, fpm)
    g = torch.Generator().manual_seed(9)
    targets = [torch.randint(0, c, (B, T), generator=g) for c in ncls]
    # ---- GPU
    out = m(wav.to(dev), target_list=[t.clone() for t in targets], padding_mask=pmask, mask=True, mask_indices=mask)
    loss, ssz, log = m.criterion(out, pred_masked_weight=1.0, pred_nomask_weight=0.5)
    loss.backward()
    torch.cuda.synchronize()
    # ---- oracle
    r = O.extract_features(sd, wav, base, padding_mask=pmask, mask_indices=mask, predict_layers=layers)
    hs = [h.transpose(0, 1) for h in r["layer_results"]]
    assert len(hs) == len(layers)
    if pre_ln:
        hs = [F.layer_norm(h, (h.shape[-1],), sd[f"post_layer_norm.{i}.weight"], sd[f"post_layer_norm.{i}.bias"], 1e-5)
              for i, h in enumerate(hs)]
    lm, lu = [], []
    sel_m, sel_u = (~fpm) & mask, (~fpm) & (~mask)
    for i, h in enumerate(hs):
        w = sd[f"final_proj.{i}.weight"] if separate else sd["final_proj.weight"]
        b = sd[f"final_proj.{i}.bias"] if separate else sd["final_proj.bias"]
        emb = sd["label_embs_concat"][i if separate else 0]
        lm += O.masked_prediction_logits(h, sel_m, targets, w, b, emb, ncls, True, 0.1)
        lu += O.masked_prediction_logits(h, sel_u, targets, w, b, emb, ncls, True, 0.1)
    want, want_ssz, _ = O.wavlm_criterion(lm, lu, 1.0, 0.5)
    want.backward()
    assert ssz == want_ssz == int(sel_m.sum()) + int(sel_u.sum())
    assert abs(float(loss) - float(want)) <= 0.02 * abs(float(want)), (float(loss), float(want))
    assert len([k for k in log if k.startswith("loss_m_")]) == len(layers) * len(ncls)
    # ---- gradients: the per-layer heads, their LayerNorms, and an encoder weight that both heads reach
    names = ["label_embs_concat", "encoder.layers.0.fc1.weight", "encoder.layers.2.fc2.weight", "post_extract_proj.weight"]
    names += [f"final_proj.{i}.weight" for i in range(len(layers))] if separate else ["final_proj.weight"]
    if pre_ln:
        names += [f"post_layer_norm.{i}.weight" for i in range(len(layers))]
    params = dict(m.named_parameters())
    for n in names:
        got, ref = params[n].grad.detach().cpu(), sd[n].grad
        assert ref is not None and got.shape == ref.shape, n
        c = _cos(got, ref)
        assert c > 0.99, (n, c)
        assert abs(float(got.norm()) - float(ref.norm())) <= 0.06 * float(ref.norm()), (n, float(got.norm()), float(ref.norm()))


def test_ils_hubert_logits_surface(cuda_device):
    """`get_logits` / `get_targets` (ils_hubert.py:290-304): layer-major list, positives in column 0; cross entropy of the
    materialised logits equals the fused criterion."""
    from unispeech_b200.ils_hubert import ILSHubertConfig, ILSHubertModel
    dev = cuda_device
    base = O.tiny_config(pre_ln=True, encoder_layers=2, relative_position_embedding=False, gru_rel_pos=False)
    cfg = ILSHubertConfig(dict(vars(base), final_dim=64, predict_layers="[1, 2]", separate_label_embeds=True))
    torch.manual_seed(5)
    m = ILSHubertModel(cfg, [30])
    m.load_state_dict(O.deterministic_state_dict(base), strict=False)
    m = m.to(dev).train()
    B, L = 2, 12000
    wav, _ = O.deterministic_waveform(B, L, seed=6)
    T = O.num_frames(L, base)
    np.random.seed(1)
    mask = m.apply_mask(B, T, None)
    targets = [torch.randint(0, 30, (B, T), generator=torch.Generator().manual_seed(2))]
    out = m(wav.to(dev), target_list=targets, padding_mask=None, mask=True, mask_indices=mask)
    loss, ssz, _ = m.criterion(out, pred_masked_weight=1.0, pred_nomask_weight=0.0)
    lg = m.get_logits(out, True)
    tg = m.get_targets(out, True)
    assert len(lg) == 2 and all(l.shape == (int(mask.sum()), 31) for l in lg) and all(int(t.sum()) == 0 for t in tg)
    ce = sum(F.cross_entropy(l, t, reduction="sum") for l, t in zip(lg, tg))
    torch.cuda.synchronize()
    assert abs(float(ce) - float(loss)) <= 0.02 * abs(float(loss)), (float(ce), float(loss))
    assert ssz == int(mask.sum())
