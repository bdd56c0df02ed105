"""CPU oracle for the WavLM / UniSpeech-SAT encoder hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional fp32 restatement (plain PyTorch on CPU) of the reference algorithm in
`/root/reference/WavLM/WavLM.py` and `/root/reference/WavLM/modules.py` (microsoft/UniSpeech @ 40d3227).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import it;
the product path (`unispeech_b200/`) never does.

Parity status: the reference ships no tests or golden vectors for this path, so the oracle is
pinned against the reference ITSELF: `tools/make_golden.py` imports the unmodified reference modules in the authoring
container, loads the deterministic parameters defined below, and commits the outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against those fixtures.

Every function cites the reference lines it follows.  Parameters are a dict keyed exactly like the reference
`state_dict`.
"""
from __future__ import annotations

import math
import zlib
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------------------------
# configuration (attribute bag mirroring WavLMConfig, WavLM/WavLM.py:162-217)
# ----------------------------------------------------------------------------------------------------------------
def make_config(**kw) -> SimpleNamespace:
    cfg = dict(
        extractor_mode="default",
        encoder_layers=12,
        encoder_embed_dim=768,
        encoder_ffn_embed_dim=3072,
        encoder_attention_heads=12,
        activation_fn="gelu",
        layer_norm_first=False,
        conv_feature_layers="[(512,10,5)] + [(512,3,2)] * 4 + [(512,2,2)] * 2",
        conv_bias=False,
        feature_grad_mult=1.0,
        normalize=False,
        dropout=0.0,
        attention_dropout=0.0,
        activation_dropout=0.0,
        encoder_layerdrop=0.0,
        dropout_input=0.0,
        dropout_features=0.0,
        mask_length=10,
        mask_prob=0.65,
        mask_selection="static",
        mask_other=0,
        no_mask_overlap=False,
        mask_min_space=1,
        mask_channel_length=10,
        mask_channel_prob=0.0,
        mask_channel_selection="static",
        mask_channel_other=0,
        no_mask_channel_overlap=False,
        mask_channel_min_space=1,
        conv_pos=128,
        conv_pos_groups=16,
        relative_position_embedding=True,
        num_buckets=320,
        max_distance=800,
        gru_rel_pos=True,
    )
    cfg.update(kw)
    return SimpleNamespace(**cfg)


def base_config(**kw) -> SimpleNamespace:
    """WavLM-Base: post-LN, GroupNorm extractor (WavLM/README.md model table; cfg of the released checkpoint)."""
    return make_config(**kw)


def large_config(**kw) -> SimpleNamespace:
    """WavLM-Large: pre-LN, LayerNorm extractor, 24 x 1024 / 4096 / 16 heads."""
    d = dict(extractor_mode="layer_norm", encoder_layers=24, encoder_embed_dim=1024, encoder_ffn_embed_dim=4096,
             encoder_attention_heads=16, layer_norm_first=True, normalize=True)
    d.update(kw)
    return make_config(**d)


def tiny_config(pre_ln: bool = False, **kw) -> SimpleNamespace:
    """Small shapes with the same structure (head_dim 64), for fixtures that fit in the repo."""
    d = dict(encoder_layers=2, encoder_embed_dim=128, encoder_ffn_embed_dim=256, encoder_attention_heads=2,
             conv_feature_layers="[(64,10,5)] + [(64,3,2)] * 4 + [(64,2,2)] * 2",
             extractor_mode="layer_norm" if pre_ln else "default", layer_norm_first=pre_ln)
    d.update(kw)
    return make_config(**d)


def conv_layers_of(cfg) -> List[Tuple[int, int, int]]:
    return eval(cfg.conv_feature_layers)  # same as WavLM/WavLM.py:230


def num_frames(L: int, cfg) -> int:
    """T_out = (T_in - k)//s + 1 per layer, no padding (nn.Conv1d defaults, WavLM/WavLM.py:400-403)."""
    for (_, k, s) in conv_layers_of(cfg):
        L = (L - k) // s + 1
    return L


# ----------------------------------------------------------------------------------------------------------------
# deterministic parameters / inputs (integer hash -> bit-identical everywhere, no RNG state involved)
# ----------------------------------------------------------------------------------------------------------------
def hash_uniform(tag: str, shape, lo: float = -1.0, hi: float = 1.0) -> Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    idx = np.arange(n, dtype=np.uint64)
    key = np.uint64(zlib.crc32(tag.encode()) & 0xFFFFFFFF)
    x = (idx * np.uint64(2654435761) + key * np.uint64(40503) + np.uint64(12345)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x45D9F3B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x45D9F3B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    u = x.astype(np.float64) / 4294967296.0
    v = (lo + (hi - lo) * u).astype(np.float32).reshape(shape)
    return torch.from_numpy(v.copy())


def parameter_shapes(cfg) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape, exactly the reference state_dict layout (WavLM/WavLM.py:220-269,378-449,507-562)."""
    D, Fd, H = cfg.encoder_embed_dim, cfg.encoder_ffn_embed_dim, cfg.encoder_attention_heads
    convs = conv_layers_of(cfg)
    C = convs[-1][0]
    sh: Dict[str, Tuple[int, ...]] = {"mask_emb": (D,)}
    cin = 1
    for i, (dim, k, _s) in enumerate(convs):
        sh[f"feature_extractor.conv_layers.{i}.0.weight"] = (dim, cin, k)
        if cfg.conv_bias:
            sh[f"feature_extractor.conv_layers.{i}.0.bias"] = (dim,)
        if cfg.extractor_mode == "layer_norm":
            sh[f"feature_extractor.conv_layers.{i}.2.1.weight"] = (dim,)
            sh[f"feature_extractor.conv_layers.{i}.2.1.bias"] = (dim,)
        elif i == 0:
            sh[f"feature_extractor.conv_layers.{i}.2.weight"] = (dim,)
            sh[f"feature_extractor.conv_layers.{i}.2.bias"] = (dim,)
        cin = dim
    sh["layer_norm.weight"] = (C,)
    sh["layer_norm.bias"] = (C,)
    if C != D:
        sh["post_extract_proj.weight"] = (D, C)
        sh["post_extract_proj.bias"] = (D,)
    sh["encoder.pos_conv.0.bias"] = (D,)
    sh["encoder.pos_conv.0.weight_g"] = (1, 1, cfg.conv_pos)
    sh["encoder.pos_conv.0.weight_v"] = (D, D // cfg.conv_pos_groups, cfg.conv_pos)
    for i in range(cfg.encoder_layers):
        p = f"encoder.layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            sh[p + f"self_attn.{n}.weight"] = (D, D)
            sh[p + f"self_attn.{n}.bias"] = (D,)
        if cfg.gru_rel_pos:
            sh[p + "self_attn.grep_linear.weight"] = (8, D // H)
            sh[p + "self_attn.grep_linear.bias"] = (8,)
            sh[p + "self_attn.grep_a"] = (1, H, 1, 1)
        if cfg.relative_position_embedding and i == 0:
            sh[p + "self_attn.relative_attention_bias.weight"] = (cfg.num_buckets, H)
        sh[p + "self_attn_layer_norm.weight"] = (D,)
        sh[p + "self_attn_layer_norm.bias"] = (D,)
        sh[p + "fc1.weight"] = (Fd, D)
        sh[p + "fc1.bias"] = (Fd,)
        sh[p + "fc2.weight"] = (D, Fd)
        sh[p + "fc2.bias"] = (D,)
        sh[p + "final_layer_norm.weight"] = (D,)
        sh[p + "final_layer_norm.bias"] = (D,)
    sh["encoder.layer_norm.weight"] = (D,)
    sh["encoder.layer_norm.bias"] = (D,)
    if cfg.layer_norm_first and getattr(cfg, "layer_norm_for_extract", False):  # UniSpeech-SAT encoder, unispeech_sat.py:1196-1197
        sh["encoder.layer_norm_for_extract.weight"] = (D,)
        sh["encoder.layer_norm_for_extract.bias"] = (D,)
    return sh


def deterministic_state_dict(cfg, seed: int = 0) -> Dict[str, Tensor]:
    """Hash-generated parameters with sensible magnitudes (not the reference initialiser; parity only needs the
    SAME weights on both sides)."""
    sd: Dict[str, Tensor] = {}
    for key, shape in parameter_shapes(cfg).items():
        tag = f"{seed}:{key}"
        if key.endswith("weight_g"):
            sd[key] = hash_uniform(tag, shape, 0.5, 1.5)
        elif key.endswith("grep_a"):
            sd[key] = hash_uniform(tag, shape, 0.8, 1.2)
        elif key == "mask_emb":
            sd[key] = hash_uniform(tag, shape, 0.0, 1.0)
        elif key.endswith("relative_attention_bias.weight"):
            sd[key] = hash_uniform(tag, shape, -1.0, 1.0)
        elif "layer_norm" in key or ".2.weight" in key or ".2.bias" in key or ".2.1." in key:
            sd[key] = hash_uniform(tag, shape, 0.8, 1.2) if key.endswith("weight") else hash_uniform(tag, shape, -0.1, 0.1)
        elif key.endswith("bias"):
            sd[key] = hash_uniform(tag, shape, -0.1, 0.1)
        else:  # linear / conv weights: uniform with unit-ish gain
            fan_in = int(np.prod(shape[1:]))
            a = math.sqrt(3.0 / fan_in) * (1.6 if "feature_extractor" in key else 1.0)
            sd[key] = hash_uniform(tag, shape, -a, a)
    return sd


def deterministic_waveform(B: int, L: int, seed: int = 0, lengths: Optional[List[int]] = None):
    """Synthetic 16 kHz batch in [-1,1); padded samples are exactly 0 and flagged in the mask."""
    wav = hash_uniform(f"wav:{seed}", (B, L))
    mask = torch.zeros(B, L, dtype=torch.bool)
    if lengths is not None:
        for b, n in enumerate(lengths):
            wav[b, n:] = 0.0
            mask[b, n:] = True
    return wav, mask


# ----------------------------------------------------------------------------------------------------------------
# training-mode dropout with the product's counter-based masks
# ----------------------------------------------------------------------------------------------------------------
_M32 = np.uint64(0xFFFFFFFF)
_M64 = (1 << 64) - 1


def _fmix32(h: np.ndarray) -> np.ndarray:
    """MurmurHash3 finaliser on uint64 arrays holding 32-bit values (unispeech_b200/csrc/dropout.cuh: fmix32)."""
    h = h ^ (h >> np.uint64(16))
    h = (h * np.uint64(0x85EBCA6B)) & _M32
    h = h ^ (h >> np.uint64(13))
    h = (h * np.uint64(0xC2B2AE35)) & _M32
    return h ^ (h >> np.uint64(16))


def _drop_bits(k0, k1, ctr: np.ndarray) -> np.ndarray:
    return _fmix32((((ctr ^ k1) * np.uint64(0x9E3779B1)) + k0) & _M32)


def drop_threshold16(p: float) -> int:
    return int(min(max(float(np.float32(p)) * 65536.0 + 0.5, 0.0), 65535.0))


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


class HashDropout:
    """The reference draws its dropout masks from torch's Philox streams (nn.Dropout / F.dropout, WavLM/WavLM.py:350,584,
    659-661,702-738; dropout_p of F.multi_head_attention_forward, WavLM/modules.py:551), which no other implementation can
    reproduce bit for bit.  What CAN be held to the reference is the semantics y = x * keep / (1 - p) at every site: this class
    restates (numpy, independent of the CUDA code) the product's counter-based keep/drop decisions so that the oracle applies
    exactly the masks the kernels apply, and parity with dropout enabled is then a deterministic comparison.
    Sites: 0 = dropout_input, 1 = encoder-level dropout, 4 + 4*layer + {0: dropout1, 1: activation dropout, 2: dropout3,
    3: attention probabilities} (unispeech_b200/dropout.py)."""

    SITE_INPUT, SITE_ENCODER = 0, 1

    def __init__(self, seed: int):
        self.seed = seed & _M64

    @staticmethod
    def layer_site(layer: int, which: int) -> int:
        return 4 + 4 * layer + which

    def key(self, site: int):
        z = _splitmix64((self.seed ^ (site * 0xD1342543DE82EF95)) & _M64)
        return np.uint64(z & 0xFFFFFFFF), np.uint64(z >> 32)

    def keep_rows(self, site: int, rows: int, N: int, p: float) -> np.ndarray:
        """bool [rows, N]: element (row, c) uses half (c & 1) of bits((row*N + c) >> 1)."""
        assert N % 2 == 0
        k0, k1 = self.key(site)
        bits = _drop_bits(k0, k1, np.arange(rows * N // 2, dtype=np.uint64))
        thr = np.uint64(drop_threshold16(p))
        keep = np.stack([(bits & np.uint64(0xFFFF)) >= thr, (bits >> np.uint64(16)) >= thr], axis=1)
        return keep.reshape(rows, N)

    def keep_attn(self, site: int, B: int, H: int, T: int, p: float) -> np.ndarray:
        """bool [B,H,T,T]: row id (b*H+h)*T+i gets its own key pair, key column j uses half (j & 1) of bits(j >> 1)."""
        k0, k1 = self.key(site)
        rowid = np.arange(B * H * T, dtype=np.uint64)
        rk0 = _fmix32((k0 + rowid * np.uint64(0x9E3779B1)) & _M32)[:, None]
        rk1 = _fmix32(k1 ^ ((rowid * np.uint64(0x85EBCA6B)) & _M32))[:, None]
        j2 = np.arange((T + 1) // 2, dtype=np.uint64)[None, :]
        bits = _fmix32((((j2 ^ rk1) * np.uint64(0x9E3779B1)) + rk0) & _M32)
        thr = np.uint64(drop_threshold16(p))
        keep = np.stack([(bits & np.uint64(0xFFFF)) >= thr, (bits >> np.uint64(16)) >= thr], axis=2)
        return keep.reshape(B * H * T, -1)[:, :T].reshape(B, H, T, T)

    def rows_btc(self, site: int, x: Tensor, p: float) -> Tensor:
        """F.dropout on a [B,T,C] tensor (logical row b*T + t)."""
        if p <= 0:
            return x
        B, T, C = x.shape
        keep = torch.from_numpy(self.keep_rows(site, B * T, C, p)).view(B, T, C)
        return x * keep.to(x.dtype) / (1.0 - p)

    def rows_tbc(self, site: int, x: Tensor, p: float) -> Tensor:
        """Same for the reference's T x B x C layout (the kernels index rows batch-major)."""
        if p <= 0:
            return x
        return self.rows_btc(site, x.transpose(0, 1), p).transpose(0, 1)

    def attn(self, site: int, probs: Tensor, B: int, H: int, p: float) -> Tensor:
        """Dropout on the softmax probabilities [B*H,T,T]."""
        if p <= 0:
            return probs
        T = probs.shape[-1]
        keep = torch.from_numpy(self.keep_attn(site, B, H, T, p)).view(B * H, T, T)
        return probs * keep.to(probs.dtype) / (1.0 - p)


# ----------------------------------------------------------------------------------------------------------------
# conv feature extractor  (ConvFeatureExtractionModel, WavLM/WavLM.py:378-449 ctor, 485-504 forward)
# ----------------------------------------------------------------------------------------------------------------
def conv_feature_extractor(sd: Dict[str, Tensor], source: Tensor, cfg, return_all: bool = False):
    x = source.unsqueeze(1)  # BxT -> BxCxT, WavLM.py:488
    outs = []
    for i, (dim, k, s) in enumerate(conv_layers_of(cfg)):
        w = sd[f"feature_extractor.conv_layers.{i}.0.weight"]
        b = sd.get(f"feature_extractor.conv_layers.{i}.0.bias")
        x = F.conv1d(x, w, b, stride=s)  # WavLM.py:400-403 (no padding)
        if cfg.extractor_mode == "layer_norm":  # WavLM.py:409-419: TransposeLast, Fp32LayerNorm, TransposeLast
            x = F.layer_norm(x.transpose(1, 2).float(), (dim,),
                             sd[f"feature_extractor.conv_layers.{i}.2.1.weight"].float(),
                             sd[f"feature_extractor.conv_layers.{i}.2.1.bias"].float(), 1e-5).transpose(1, 2)
        elif i == 0:  # WavLM.py:420-426: Fp32GroupNorm(dim, dim) == per-(b,c) statistics over time
            x = F.group_norm(x.float(), dim, sd[f"feature_extractor.conv_layers.{i}.2.weight"].float(),
                             sd[f"feature_extractor.conv_layers.{i}.2.bias"].float(), 1e-5)
        x = F.gelu(x)  # nn.GELU(): exact erf form
        outs.append(x)
    return (x, outs) if return_all else x


def frame_padding_mask(padding_mask: Tensor, T: int) -> Tensor:
    """WavLM.forward_padding_mask, WavLM/WavLM.py:311-321."""
    extra = padding_mask.size(1) % T
    if extra > 0:
        padding_mask = padding_mask[:, :-extra]
    return padding_mask.view(padding_mask.size(0), T, -1).all(-1)


# ----------------------------------------------------------------------------------------------------------------
# relative position bias  (MultiheadAttention._relative_positions_bucket / compute_bias, WavLM/modules.py:417-455)
# ----------------------------------------------------------------------------------------------------------------
def relative_positions_bucket(relative_positions: Tensor, num_buckets: int, max_distance: int) -> Tensor:
    nb = num_buckets // 2  # bidirectional
    buckets = (relative_positions > 0).to(torch.long) * nb
    rp = torch.abs(relative_positions)
    max_exact = nb // 2
    is_small = rp < max_exact
    large = max_exact + (
        torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)
    ).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(is_small, rp, large)


def bucket_lut(T: int, num_buckets: int, max_distance: int) -> Tensor:
    """bucket(delta) for delta = j - i in [-(T-1), T-1]; index delta + T - 1.  The bias matrix is Toeplitz."""
    delta = torch.arange(-(T - 1), T, dtype=torch.long)
    return relative_positions_bucket(delta, num_buckets, max_distance)


def compute_bias(T: int, emb_weight: Tensor, num_buckets: int, max_distance: int) -> Tensor:
    """[H, T, T] position bias exactly as modules.py:445-455 (dense form)."""
    ctx = torch.arange(T, dtype=torch.long)[:, None]
    mem = torch.arange(T, dtype=torch.long)[None, :]
    bucket = relative_positions_bucket(mem - ctx, num_buckets, max_distance)
    return F.embedding(bucket, emb_weight).permute(2, 0, 1)


# ----------------------------------------------------------------------------------------------------------------
# attention  (MultiheadAttention.forward fast path, WavLM/modules.py:457-564)
# ----------------------------------------------------------------------------------------------------------------
def gate_values(x_tbc: Tensor, grep_w: Tensor, grep_b: Tensor, grep_a: Tensor, H: int) -> Tensor:
    """gru_rel_pos gate on the RAW layer input (modules.py:523-533).  Returns [B,H,T,1]."""
    T, B, D = x_tbc.shape
    q = x_tbc.transpose(0, 1).view(B, T, H, D // H).permute(0, 2, 1, 3)
    g = torch.sigmoid(F.linear(q, grep_w, grep_b).view(B, H, T, 2, 4).sum(-1))
    gate_a, gate_b = g.chunk(2, dim=-1)
    return gate_a * (gate_b * grep_a - 1.0) + 2.0


def self_attention(sd, prefix: str, x_tbc: Tensor, key_padding_mask: Optional[Tensor], position_bias: Optional[Tensor],
                   cfg, drop: Optional["HashDropout"] = None, drop_site: int = 0) -> Tensor:
    """softmax((xWq+bq)/sqrt(d) (xWk+bk)^T + gate*bias, -inf at padded keys) (xWv+bv), then out_proj."""
    T, B, D = x_tbc.shape
    H = cfg.encoder_attention_heads
    hd = D // H
    q = F.linear(x_tbc, sd[prefix + "q_proj.weight"], sd[prefix + "q_proj.bias"]) * (hd ** -0.5)
    k = F.linear(x_tbc, sd[prefix + "k_proj.weight"], sd[prefix + "k_proj.bias"])
    v = F.linear(x_tbc, sd[prefix + "v_proj.weight"], sd[prefix + "v_proj.bias"])
    q = q.contiguous().view(T, B * H, hd).transpose(0, 1)
    k = k.contiguous().view(T, B * H, hd).transpose(0, 1)
    v = v.contiguous().view(T, B * H, hd).transpose(0, 1)
    scores = torch.bmm(q, k.transpose(1, 2))  # [B*H, T, T]
    if position_bias is not None:
        bias = position_bias  # [B*H, T, T]
        if cfg.gru_rel_pos:
            g = gate_values(x_tbc, sd[prefix + "grep_linear.weight"], sd[prefix + "grep_linear.bias"],
                            sd[prefix + "grep_a"], H)
            bias = g.view(B * H, T, 1) * position_bias
        scores = scores + bias
    if key_padding_mask is not None:
        scores = scores.view(B, H, T, T).masked_fill(key_padding_mask[:, None, None, :], float("-inf")).view(B * H, T, T)
    p = torch.softmax(scores, dim=-1)
    if drop is not None:  # dropout_p of F.multi_head_attention_forward (modules.py:551): dropout on the probabilities
        p = drop.attn(drop_site, p, B, H, cfg.attention_dropout)
    o = torch.bmm(p, v).transpose(0, 1).contiguous().view(T, B, D)
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


# ----------------------------------------------------------------------------------------------------------------
# encoder layer / stack  (TransformerSentenceEncoderLayer.forward WavLM/WavLM.py:677-742; TransformerEncoder :564-612)
# ----------------------------------------------------------------------------------------------------------------
def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def encoder_layer(sd, i: int, x: Tensor, padding_mask, position_bias, cfg, drop: Optional["HashDropout"] = None) -> Tensor:
    p = f"encoder.layers.{i}."
    site = lambda which: HashDropout.layer_site(i, which)
    d1 = (lambda t: drop.rows_tbc(site(0), t, cfg.dropout)) if drop is not None else (lambda t: t)  # self.dropout1, :659
    d2 = (lambda t: drop.rows_tbc(site(1), t, cfg.activation_dropout)) if drop is not None else (lambda t: t)  # dropout2, :660
    d3 = (lambda t: drop.rows_tbc(site(2), t, cfg.dropout)) if drop is not None else (lambda t: t)  # self.dropout3, :661
    residual = x
    if cfg.layer_norm_first:  # WavLM.py:691-714
        x = _ln(x, sd, p + "self_attn_layer_norm")
        x = self_attention(sd, p + "self_attn.", x, padding_mask, position_bias, cfg, drop, site(3))
        x = residual + d1(x)
        residual = x
        x = _ln(x, sd, p + "final_layer_norm")
        x = d2(F.gelu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"])))
        x = F.linear(x, sd[p + "fc2.weight"], sd[p + "fc2.bias"])
        x = residual + d3(x)
    else:  # WavLM.py:715-740
        x = self_attention(sd, p + "self_attn.", x, padding_mask, position_bias, cfg, drop, site(3))
        x = residual + d1(x)
        x = _ln(x, sd, p + "self_attn_layer_norm")
        residual = x
        x = d2(F.gelu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"])))
        x = F.linear(x, sd[p + "fc2.weight"], sd[p + "fc2.bias"])
        x = residual + d3(x)
        x = _ln(x, sd, p + "final_layer_norm")
    return x


def pos_conv_weight(sd) -> Tensor:
    """nn.utils.weight_norm(dim=2): w = g * v / ||v|| with the norm over dims (0,1) per tap."""
    v = sd["encoder.pos_conv.0.weight_v"]
    g = sd["encoder.pos_conv.0.weight_g"]
    return g * v / v.norm(2, dim=(0, 1), keepdim=True)


def encoder(sd, x: Tensor, padding_mask: Optional[Tensor], cfg, tgt_layer=None, extract_layer: Optional[int] = None,
            drop: Optional["HashDropout"] = None):
    """TransformerEncoder.forward + extract_features (out-of-place restatement of WavLM.py:564-612).
    Variants of the fairseq tree: `tgt_layer` may be a LIST of 1-based layer numbers (src/fairseq/models/wavlm/wavlm.py:
    730-737: their outputs are collected, there is no pre-layer entry and no early exit); `extract_layer` (0-based) returns
    that layer's output as a third value, passed through `encoder.layer_norm_for_extract` for pre-LN models when no target
    layer is set (src/fairseq/models/unispeech_sat/unispeech_sat.py:1202-1255)."""
    if padding_mask is not None:
        x = x.masked_fill(padding_mask.unsqueeze(-1), 0.0)  # x[padding_mask] = 0, :574-575
    w = pos_conv_weight(sd)
    x_conv = F.conv1d(x.transpose(1, 2), w, sd["encoder.pos_conv.0.bias"], padding=cfg.conv_pos // 2,
                      groups=cfg.conv_pos_groups)
    if cfg.conv_pos % 2 == 0:
        x_conv = x_conv[:, :, :-1]  # SamePad, modules.py:72-83
    x = x + F.gelu(x_conv).transpose(1, 2)  # :577-579
    if not cfg.layer_norm_first:
        x = _ln(x, sd, "encoder.layer_norm")  # :581-582
    if drop is not None:
        x = drop.rows_btc(HashDropout.SITE_ENCODER, x, cfg.dropout)  # x = F.dropout(x, p=self.dropout, training), :584
    x = x.transpose(0, 1)  # B x T x C -> T x B x C
    layer_results = []
    tgt_list = tgt_layer if isinstance(tgt_layer, (list, tuple)) else None
    if tgt_layer is not None and tgt_list is None:
        layer_results.append(x)
    T, B, D = x.shape
    H = cfg.encoder_attention_heads
    position_bias = None
    if cfg.relative_position_embedding:
        pb = compute_bias(T, sd["encoder.layers.0.self_attn.relative_attention_bias.weight"], cfg.num_buckets,
                          cfg.max_distance)
        position_bias = pb.unsqueeze(0).repeat(B, 1, 1, 1).view(B * H, T, T)  # modules.py:504-506
    r = None
    er = None
    for i in range(cfg.encoder_layers):
        x = encoder_layer(sd, i, x, padding_mask, position_bias, cfg, drop)
        if tgt_list is not None:
            if i + 1 in tgt_list:
                layer_results.append(x)
        elif tgt_layer is not None:
            layer_results.append(x)
        if extract_layer is not None and i == extract_layer:
            er = x.transpose(0, 1)
        if tgt_list is None and i == tgt_layer:
            r = x
            break
    if r is not None:
        x = r
    x = x.transpose(0, 1)
    if cfg.layer_norm_first and tgt_layer is None:  # :567-568
        x = _ln(x, sd, "encoder.layer_norm")
        if er is not None and "encoder.layer_norm_for_extract.weight" in sd:
            er = _ln(er, sd, "encoder.layer_norm_for_extract")
    if extract_layer is not None:
        return x, layer_results, er
    return x, layer_results


def extract_features(sd, source: Tensor, cfg, padding_mask: Optional[Tensor] = None,
                     mask_indices: Optional[Tensor] = None, output_layer: Optional[int] = None,
                     drop: Optional["HashDropout"] = None, predict_layers: Optional[List[int]] = None):
    """WavLM.extract_features, WavLM/WavLM.py:323-375.  `mask_indices` [B,T] bool replaces the host-RNG
    compute_mask_indices call of apply_mask (:271-309): masked frames are set to mask_emb."""
    feats = conv_feature_extractor(sd, source, cfg)  # :333-339 (feature_grad_mult handled by callers)
    feats = feats.transpose(1, 2)
    C = feats.shape[-1]
    feats = F.layer_norm(feats, (C,), sd["layer_norm.weight"], sd["layer_norm.bias"], 1e-5)  # :342
    if padding_mask is not None:
        padding_mask = frame_padding_mask(padding_mask, feats.size(1))
    if "post_extract_proj.weight" in sd:
        feats = F.linear(feats, sd["post_extract_proj.weight"], sd["post_extract_proj.bias"])
    if drop is not None:
        feats = drop.rows_btc(HashDropout.SITE_INPUT, feats, cfg.dropout_input)  # features = self.dropout_input(features), :350
    x = feats
    if mask_indices is not None:
        x = torch.where(mask_indices.unsqueeze(-1), sd["mask_emb"].to(x.dtype), x)  # x[mask_indices] = mask_emb
    # `predict_layers` (1-based list): ILS-HuBERT's `layer=self.predict_layers` (src/fairseq/models/hubert/ils_hubert.py:167-171)
    tgt = list(predict_layers) if (predict_layers is not None and output_layer is None) else \
        (None if output_layer is None else output_layer - 1)
    x, layer_results = encoder(sd, x, padding_mask, cfg, tgt, drop=drop)
    return {"x": x, "padding_mask": padding_mask, "features": feats, "layer_results": layer_results}


def probe_loss(x: Tensor, padding_mask: Optional[Tensor], seed: int = 0) -> Tensor:
    """Non-degenerate scalar for gradient parity: sum(x * R) over valid frames, R hash-generated."""
    R = hash_uniform(f"probe:{seed}", tuple(x.shape)).to(x.dtype).to(x.device)
    if padding_mask is not None:
        R = R.masked_fill(padding_mask.unsqueeze(-1).to(R.device), 0.0)
    return (x * R).sum()


def forward_flops(L: int, cfg) -> float:
    """Algorithmic GEMM flops of one forward pass of one utterance (SURVEY.md section 8d formula)."""
    D, Fd, H = cfg.encoder_embed_dim, cfg.encoder_ffn_embed_dim, cfg.encoder_attention_heads
    fl, cin, t = 0.0, 1, L
    for (dim, k, s) in conv_layers_of(cfg):
        t = (t - k) // s + 1
        fl += 2.0 * t * dim * cin * k
        cin = dim
    T = t
    fl += 2.0 * T * cin * D
    fl += 2.0 * T * D * (D // cfg.conv_pos_groups) * cfg.conv_pos
    fl += cfg.encoder_layers * (8.0 * T * D * D + 4.0 * T * T * D + 4.0 * T * D * Fd + 2.0 * T * H * 64 * 8)
    return fl


# ----------------------------------------------------------------------------------------------------------------
# optimizer step  (utils.clip_grad_norm_ src/fairseq/utils.py:338-381; FP16Optimizer.multiply_grads / clip_grad_norm
# src/fairseq/optim/fp16_optimizer.py:190-214; Adam.step src/fairseq/optim/adam.py:150-228)
# ----------------------------------------------------------------------------------------------------------------
def clip_coefficient(grads: List[Tensor], max_norm: float, multiply_factor: float = 1.0):
    """Returns (grad_norm, factor to multiply every gradient with): norm of the per-tensor norms (utils.py:359-375) times the
    pending multiply factor (fp16_optimizer.py:198-200), clip_coef = (max_norm / (norm + 1e-6)).clamp(max=1) (:207-209)."""
    total = torch.norm(torch.stack([torch.norm(g.float(), p=2) for g in grads])) * abs(multiply_factor)
    coef = multiply_factor
    if max_norm > 0:
        coef = coef * float((max_norm / (total + 1e-6)).clamp(max=1.0))
    return total, coef


def adam_step(params: List[Tensor], grads: List[Tensor], state: dict, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
              weight_decay: float = 0.0):
    """In-place fairseq Adam update of fp32 tensors (adam.py:176-222, amsgrad off); `state` holds step / exp_avg / exp_avg_sq."""
    if not state:
        state.update(step=0, exp_avg=[torch.zeros_like(p) for p in params], exp_avg_sq=[torch.zeros_like(p) for p in params])
    state["step"] += 1
    beta1, beta2 = betas
    bias_correction1 = 1 - beta1 ** state["step"]
    bias_correction2 = 1 - beta2 ** state["step"]
    step_size = lr * math.sqrt(bias_correction2) / bias_correction1
    for p, g, m, v in zip(params, grads, state["exp_avg"], state["exp_avg_sq"]):
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = v.sqrt().add_(eps)
        if weight_decay != 0:
            p.add_(p, alpha=-weight_decay * lr)
        p.addcdiv_(m, denom, value=-step_size)


# ----------------------------------------------------------------------------------------------------------------
# masked-prediction head + criterion  (src/fairseq/models/wavlm/wavlm.py:426-438 compute_nce, :525-576 forward tail;
# src/fairseq/criterions/wavlm_criterion.py:52-138 get_loss)
# ----------------------------------------------------------------------------------------------------------------
def compute_nce(x: Tensor, pos: Tensor, negs: Tensor, logit_temp: float) -> Tensor:
    """wavlm.py:426-438 verbatim in structure: [S, C+1] logits, column 0 = positive, -inf where a negative equals the positive."""
    neg_is_pos = (pos == negs).all(-1)
    pos = pos.unsqueeze(0)
    targets = torch.cat([pos, negs], dim=0)
    logits = torch.cosine_similarity(x.float(), targets.float(), dim=-1).type_as(x)
    logits = logits / logit_temp
    if neg_is_pos.any():
        logits[1:][neg_is_pos] = float("-inf")
    return logits.transpose(0, 1)


def masked_prediction_logits(x: Tensor, sel: Tensor, target_list: List[Tensor], final_proj_w: Tensor, final_proj_b: Tensor,
                             label_embs_concat: Tensor, num_classes: List[int], untie_final_proj: bool, logit_temp: float):
    """Logit lists of wavlm.py:525-553 for the frames selected by the bool [B,T] mask `sel` (masked or unmasked set)."""
    proj = F.linear(x[sel], final_proj_w, final_proj_b)
    projs = proj.chunk(len(target_list), dim=-1) if untie_final_proj else [proj for _ in target_list]
    label_embs_list = label_embs_concat.split(num_classes, 0)
    out = []
    for i, (p, t) in enumerate(zip(projs, target_list)):
        y = torch.index_select(label_embs_list[i], 0, t[sel].long())
        negs = label_embs_list[i].unsqueeze(1).expand(-1, p.size(0), -1)
        out.append(compute_nce(p, y, negs, logit_temp))
    return out


def wavlm_criterion(logit_m_list: List[Tensor], logit_u_list: List[Tensor], pred_masked_weight: float, pred_nomask_weight: float,
                    features_pen: Optional[Tensor] = None, loss_weights: Optional[List[float]] = None):
    """WavLMCriterion.get_loss (wavlm_criterion.py:52-103, :116-138 for the accuracy counts): positives sit at index 0."""
    loss, sample_size, log = 0.0, 0, {}
    for tag, lst, w in (("m", logit_m_list, pred_masked_weight), ("u", logit_u_list, pred_nomask_weight)):
        parts = []
        for i, lg in enumerate(lst):
            tgt = lg.new_zeros(lg.size(0), dtype=torch.long)
            l_ = F.cross_entropy(lg.float(), tgt, reduction="sum")
            parts.append(l_)
            log[f"loss_{tag}_{i}"] = l_.detach()
            mx, mn = lg.argmax(-1) == 0, lg.argmin(-1) == 0
            log[f"correct_{tag}_{i}"] = int(mx.long().sum() - (mx & mn).long().sum())
            log[f"count_{tag}_{i}"] = mx.numel()
        if w > 0 and parts:
            loss = loss + w * sum(parts)
            sample_size += lst[0].size(0)
    if loss_weights is not None and features_pen is not None and loss_weights[0] != 0:
        loss = loss + loss_weights[0] * features_pen.float() * sample_size
    return loss, sample_size, log


# ----------------------------------------------------------------------------------------------------------------
# UniSpeech-SAT utterance-contrastive branch (BASELINE config #4; no CUDA path yet: oracle groundwork)
#   sample_instances   src/fairseq/models/unispeech_sat/unispeech_sat.py:487-543
#   compute_nce        :545-557          compute_pred_spk / forward tail   :699-745
#   GumbelVectorQuantizer.forward (eval mode: hard arg-max codes)   src/fairseq/modules/gumbel_vector_quantizer.py:141-201
# ----------------------------------------------------------------------------------------------------------------
def sat_sample_instances(y: Tensor, num: int, n_instances: int, cross_sample_instances: int):
    """Negative 'instances' for every position: `n_instances` drawn inside the utterance and `cross_sample_instances` drawn over
    the whole batch, never the position itself.  Same torch.randint call order as the reference, so it is reproducible under
    torch.manual_seed.  Returns (instances [N, B, T, C], flat indices [B, N*T])."""
    bsz, tsz, fsz = y.shape
    y = y.reshape(-1, fsz)
    cross_high, high = tsz * bsz, tsz
    with torch.no_grad():
        if n_instances > 0:
            tszs = torch.arange(num).unsqueeze(-1).expand(-1, n_instances).flatten()
            instance_idxs = torch.randint(low=0, high=high - 1, size=(bsz, n_instances * num))
            instance_idxs[instance_idxs >= tszs] += 1
        if cross_sample_instances > 0:
            tszs = torch.arange(num).unsqueeze(-1).expand(-1, cross_sample_instances).flatten()
            cross_instance_idxs = torch.randint(low=0, high=cross_high - 1, size=(bsz, cross_sample_instances * num))
            cross_instance_idxs[cross_instance_idxs >= tszs] += 1
    if n_instances > 0:
        for i in range(1, bsz):
            instance_idxs[i] += i * high
    else:
        instance_idxs = cross_instance_idxs
    if cross_sample_instances > 0 and n_instances > 0:
        instance_idxs = torch.cat([instance_idxs, cross_instance_idxs], dim=1)
    instances = y[instance_idxs.view(-1)]
    instances = instances.view(bsz, n_instances + cross_sample_instances, num, fsz).permute(1, 0, 2, 3)
    return instances, instance_idxs


def sat_compute_nce(x: Tensor, pos: Tensor, instances: Tensor, logit_temp: float, replace_inf: bool = True) -> Tensor:
    instance_is_pos = (pos == instances).all(-1)
    targets = torch.cat([pos.unsqueeze(0), instances], dim=0)
    logits = torch.cosine_similarity(x.float(), targets.float(), dim=-1).type_as(x) / logit_temp
    if instance_is_pos.any() and replace_inf:
        logits[1:][instance_is_pos] = float("-inf")
    return logits.transpose(0, 1)


def gumbel_noise(seed: int, site: int, n: int) -> Tensor:
    """Gumbel(0,1) samples 0..n-1 of the product's counter-based generator (csrc/sat.cu gumbel_noise, keys as for dropout):
    u = (float32(bits) + 0.5) * 2^-32 clamped below 1, g = -log(-log u)."""
    k0, k1 = HashDropout(seed).key(site)
    bits = _drop_bits(k0, k1, np.arange(n, dtype=np.uint64))
    u = (bits.astype(np.float32) + np.float32(0.5)) * np.float32(2.3283064365386963e-10)
    u = np.minimum(u, np.float32(0.99999994)).astype(np.float64)
    return torch.from_numpy(-np.log(-np.log(u))).float()


def gumbel_vq_eval(x: Tensor, weight_proj_w: Tensor, weight_proj_b: Tensor, vars_: Tensor, groups: int, num_vars: int,
                   noise: Optional[Tensor] = None, tau: float = 1.0):
    """GumbelVectorQuantizer.forward (time_first, combine_groups=False, weight_proj_depth=1), gumbel_vector_quantizer.py:141-201.
    Eval mode (noise None): hard arg-max code per group.  Training mode: `F.gumbel_softmax(x, tau, hard=True)` with the Gumbel
    noise supplied by the caller ([B*T*G, V]; torch draws it from its Philox stream): one-hot of the soft sample's arg-max in the
    forward value, the soft sample's gradient in the backward pass.  Plus the two perplexities the criterion logs / penalises."""
    bsz, tsz, fsz = x.shape
    lg = F.linear(x.reshape(-1, fsz), weight_proj_w, weight_proj_b).view(bsz * tsz * groups, -1)
    k = lg.argmax(-1)
    hard = lg.new_zeros(*lg.shape).scatter_(-1, k.view(-1, 1), 1.0).view(bsz * tsz, groups, -1)
    hard_probs = hard.float().mean(0)
    code_ppl = torch.exp(-torch.sum(hard_probs * torch.log(hard_probs + 1e-7), dim=-1)).sum()
    avg_probs = torch.softmax(lg.view(bsz * tsz, groups, -1).float(), dim=-1).mean(0)
    prob_ppl = torch.exp(-torch.sum(avg_probs * torch.log(avg_probs + 1e-7), dim=-1)).sum()
    if noise is not None:   # F.gumbel_softmax(hard=True): y_hard - y_soft.detach() + y_soft
        y_soft = torch.softmax((lg.float() + noise) / tau, dim=-1)
        idx = y_soft.argmax(-1, keepdim=True)
        y_hard = torch.zeros_like(y_soft).scatter_(-1, idx, 1.0)
        sel = (y_hard - y_soft.detach() + y_soft).view(bsz * tsz, groups, -1)
        k = idx.view(-1)
    else:
        sel = hard
    q = (sel.view(bsz * tsz, -1).unsqueeze(-1) * vars_).view(bsz * tsz, groups, num_vars, -1).sum(-2).view(bsz, tsz, -1)
    return {"x": q, "code_perplexity": code_ppl, "prob_perplexity": prob_ppl, "num_vars": num_vars * groups,
            "codes": k.view(bsz * tsz, groups)}   # the selected code per (frame, group)


def sat_utterance_contrastive_loss(spk_x: Tensor, padding_mask: Tensor, mask_indices: Tensor, spk_proj_w: Tensor, spk_proj_b: Tensor,
                                   n_instances: int, cross_sample_instances: int, logit_temp: float, quantizer=None,
                                   project_q=None):
    """Masked branch of unispeech_sat.py:699-758.  `spk_x` [B,T,C] is the `utterance_contrastive_layer` output (the encoder's
    `extract_layer` result); every utterance must have the same number of masked, unpadded frames (`.view(B, -1, C)`, :742).
    `quantizer` = dict(weight_proj_w, weight_proj_b, vars, groups, num_vars) in eval mode or None; `project_q` = (w, b) or None.
    Returns (loss_spk, mean_targets, contrastive_acc, quantizer outputs or None)."""
    B = spk_x.size(0)
    b_pos = torch.arange(B).unsqueeze(1).expand(B, spk_x.size(1))
    masked = ~padding_mask & mask_indices
    x = spk_x[masked].view(B, -1, spk_x.size(-1))
    proj_x = F.linear(x, spk_proj_w, spk_proj_b)
    x_b_pos = b_pos[masked].view(B, -1)
    q = None
    if quantizer is not None:
        q = gumbel_vq_eval(x, **quantizer)
        y = F.linear(q["x"], project_q[0], project_q[1])
    else:
        y = proj_x
    N = n_instances + cross_sample_instances
    samples, samples_idx = sat_sample_instances(y, y.size(1), n_instances, cross_sample_instances)
    samples_b_pos = x_b_pos.reshape(-1)[samples_idx.view(-1)].view(B, N, x.size(1)).permute(1, 0, 2)
    x_pos_batch = x_b_pos[..., 0].unsqueeze(1).unsqueeze(0).expand_as(samples_b_pos)
    samples_targets = (samples_b_pos == x_pos_batch).long()
    targets = torch.cat((x.new_ones(1, B, x.size(1), dtype=torch.long), samples_targets), dim=0)
    proj_flat = proj_x.reshape(-1, proj_x.size(-1))
    y_flat = y.reshape(-1, y.size(-1))
    samples = samples.reshape(samples.size(0), -1, y_flat.size(-1))
    targets = targets.reshape(targets.size(0), -1).transpose(0, 1)
    logits = sat_compute_nce(proj_flat, y_flat, samples, logit_temp, replace_inf=False)
    loss = F.binary_cross_entropy_with_logits(logits, targets.type_as(logits), reduction="none").mean()
    return loss, targets.float().mean(), ((logits >= 0.0) == targets).float().mean(), q


# ----------------------------------------------------------------------------------------------------------------
# wav2vec 2.0 contrastive head
#   sample_negatives   src/fairseq/models/wav2vec/wav2vec2.py:474-531   (the draws of sat_sample_instances, frame-major layout)
#   compute_preds      :533-553          forward tail :621-723          get_logits / get_targets / get_extra_losses :741-767
#   Wav2vecCriterion.get_loss (infonce)  src/fairseq/criterions/wav2vec_criterion.py:44-118
# Pinned by tests/golden/w2v_heads.npz (tools/make_w2v_golden.py executes the reference's source text).
# ----------------------------------------------------------------------------------------------------------------
def w2v_sample_negatives(y: Tensor, num: int, n_negatives: int, cross_sample_negatives: int, padding_count: int = 0):
    """wav2vec2.py:474-531: the same torch.randint draws as the UniSpeech-SAT sampler, but the flat index list of an utterance is
    read as [num, N] (frame-major: `negs.view(bsz, num, N, fsz).permute(2, 0, 1, 3)`), not [N, num].  Returns (negs N x B x T x C,
    flat indices [B, N * num])."""
    bsz, tsz, fsz = y.shape
    yf = y.reshape(-1, fsz)
    cross_high, high = tsz * bsz, tsz - padding_count
    assert high > 1
    with torch.no_grad():
        if n_negatives > 0:
            tszs = torch.arange(num).unsqueeze(-1).expand(-1, n_negatives).flatten()
            neg_idxs = torch.randint(low=0, high=high - 1, size=(bsz, n_negatives * num))
            neg_idxs[neg_idxs >= tszs] += 1
        if cross_sample_negatives > 0:
            tszs = torch.arange(num).unsqueeze(-1).expand(-1, cross_sample_negatives).flatten()
            cross_neg_idxs = torch.randint(low=0, high=cross_high - 1, size=(bsz, cross_sample_negatives * num))
            cross_neg_idxs[cross_neg_idxs >= tszs] += 1
    if n_negatives > 0:
        for i in range(1, bsz):
            neg_idxs[i] += i * high
    else:
        neg_idxs = cross_neg_idxs
    if cross_sample_negatives > 0 and n_negatives > 0:
        neg_idxs = torch.cat([neg_idxs, cross_neg_idxs], dim=1)
    negs = yf[neg_idxs.view(-1)].view(bsz, num, n_negatives + cross_sample_negatives, fsz).permute(2, 0, 1, 3)
    return negs, neg_idxs


def w2v_compute_preds(x: Tensor, y: Tensor, negatives: Tensor, logit_temp: float) -> Tensor:
    """[N+1, B, T] logits: cosine similarity of x with the positive y and the negatives, / temp; a negative that equals the
    positive is set to -inf (wav2vec2.py:533-553)."""
    neg_is_pos = (y == negatives).all(-1)
    targets = torch.cat([y.unsqueeze(0), negatives], dim=0)
    logits = torch.cosine_similarity(x.float(), targets.float(), dim=-1).type_as(x) / logit_temp
    if neg_is_pos.any():
        logits[1:][neg_is_pos] = float("-inf")
    return logits


def w2v_contrastive_loss(x_enc: Tensor, unmasked_features: Tensor, mask_indices: Tensor, final_proj, project_q, n_negatives: int,
                         cross_sample_negatives: int, logit_temp: float, quantizer=None, noise: Optional[Tensor] = None,
                         tau: float = 1.0):
    """Masked branch of Wav2Vec2Model.forward (:621-723) + the infonce criterion.  `x_enc` [B,T,D] encoder output,
    `unmasked_features` [B,T,C] LayerNorm'ed conv features, `mask_indices` bool [B,T] with the same count per utterance;
    `final_proj` / `project_q` = (w, b); `quantizer` = dict(weight_proj_w, weight_proj_b, vars_, groups, num_vars) or None.
    Returns dict(loss (sum CE), sample_size, correct, count, logits [S, N+1], q)."""
    B = x_enc.size(0)
    y = unmasked_features[mask_indices].view(B, -1, unmasked_features.size(-1))
    q = None
    if quantizer is not None:
        q = gumbel_vq_eval(y, noise=noise, tau=tau, **quantizer)
        y = F.linear(q["x"], project_q[0], project_q[1])
    else:
        y = F.linear(y, project_q[0], project_q[1])
    negs, _ = w2v_sample_negatives(y, y.size(1), n_negatives, cross_sample_negatives)   # N x B x T x C
    x = x_enc[mask_indices].view(B, -1, x_enc.size(-1))
    x = F.linear(x, final_proj[0], final_proj[1])
    logits = w2v_compute_preds(x, y, negs, logit_temp)                  # [N+1, B, T]
    lg = logits.transpose(0, 2).reshape(-1, logits.size(0)).float()     # get_logits, :741-745
    tgt = lg.new_zeros(lg.size(0), dtype=torch.long)
    loss = F.cross_entropy(lg, tgt, reduction="sum")
    mx, mn = lg.argmax(-1) == 0, lg.argmin(-1) == 0
    return {"loss": loss, "sample_size": tgt.numel(), "correct": int(mx.long().sum() - (mx & mn).long().sum()), "count": mx.numel(),
            "logits": lg, "q": q}


def w2v_criterion(head: dict, features_pen: Optional[Tensor], loss_weights: Optional[List[float]]):
    """Wav2vecCriterion.get_loss with infonce (:44-87): loss + sum_i w_i * extra_i * sample_size, extras = [(num_vars - prob_ppl) /
    num_vars, features_pen] (get_extra_losses, wav2vec2.py:752-767)."""
    loss, ssz = head["loss"], head["sample_size"]
    extras = []
    if head["q"] is not None:
        nv = head["q"]["num_vars"]
        extras.append((nv - head["q"]["prob_perplexity"]) / nv)
    if features_pen is not None:
        extras.append(features_pen)
    if loss_weights is not None:
        lw = list(loss_weights)
        if len(lw) == 1 and len(extras) != 1:
            lw = [lw[0]] * len(extras)
        assert len(lw) == len(extras)
        for p_, c in zip(extras, lw):
            if c != 0 and p_ is not None:
                loss = loss + c * p_.float() * ssz
    return loss, ssz
