"""Deterministic synthetic weights and clips for tests, smoke and bench (there is no network for real
checkpoints or LRW data).  Keys/shapes follow the reference state dict exactly, so the same
dict loads into the reference modules, the oracle and the drop-in modules.

Distributions follow SURVEY.md §8d: xavier-uniform-like matrices / conv kernels (what Transformer.__init__
leaves behind, transformer/transformer.py:18-20) and randomised BatchNorm statistics, so that the BN fold is
actually exercised (default gamma=1, beta=0, mean=0, var=1 would make it an identity).
"""
from __future__ import annotations

import math

import torch


def _xavier(gen, shape):
    """xavier_uniform_ bound for a weight of `shape` ([out, in, *kernel])."""
    rf = 1
    for s in shape[2:]:
        rf *= s
    fan_in, fan_out = shape[1] * rf, shape[0] * rf
    bound = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen) * 2.0 - 1.0) * bound


def _bn(gen, sd, prefix, c):
    sd[prefix + ".weight"] = torch.rand(c, generator=gen) + 0.5          # U(0.5,1.5)
    sd[prefix + ".bias"] = torch.randn(c, generator=gen) * 0.1
    sd[prefix + ".running_mean"] = torch.randn(c, generator=gen) * 0.1
    sd[prefix + ".running_var"] = torch.rand(c, generator=gen) + 0.5
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def frontend_state_dict(seed=1, prefix=""):
    """State dict of reference `Lipreading` (transformer/video_frontend.py:91-109)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    sd[prefix + "frontend3D.0.weight"] = _xavier(gen, (64, 1, 5, 7, 7))
    _bn(gen, sd, prefix + "frontend3D.1", 64)
    inplanes = 64
    for li, planes in ((1, 64), (2, 128), (3, 256), (4, 512)):
        for bi in range(2):
            p = f"{prefix}resnet18.layer{li}.{bi}"
            cin = inplanes if bi == 0 else planes
            sd[p + ".conv1.weight"] = _xavier(gen, (planes, cin, 3, 3))
            _bn(gen, sd, p + ".bn1", planes)
            sd[p + ".conv2.weight"] = _xavier(gen, (planes, planes, 3, 3))
            _bn(gen, sd, p + ".bn2", planes)
            if bi == 0 and li != 1:
                sd[p + ".downsample.0.weight"] = _xavier(gen, (planes, cin, 1, 1))
                _bn(gen, sd, p + ".downsample.1", planes)
        inplanes = planes
    return sd


def encoder_state_dict(seed=2, n_layers=6, d_input=512, d_model=512, d_inner=2048, n_head=8, d_k=64, d_v=64,
                       pe_maxlen=5000, prefix=""):
    """State dict of reference `Encoder` (transformer/encoder.py:12-34)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}

    def linear(name, out_f, in_f):
        sd[name + ".weight"] = _xavier(gen, (out_f, in_f))
        sd[name + ".bias"] = (torch.rand(out_f, generator=gen) * 2 - 1) / math.sqrt(in_f)

    def ln(name):
        sd[name + ".weight"] = torch.rand(d_model, generator=gen) + 0.5
        sd[name + ".bias"] = torch.randn(d_model, generator=gen) * 0.1

    linear(prefix + "linear_in", d_model, d_input)
    ln(prefix + "layer_norm_in")
    pe = torch.zeros(pe_maxlen, d_model)
    position = torch.arange(0, pe_maxlen).unsqueeze(1).float()
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    sd[prefix + "positional_encoding.pe"] = pe.unsqueeze(0)
    for i in range(n_layers):
        p = f"{prefix}layer_stack.{i}"
        linear(p + ".slf_attn.w_qs", n_head * d_k, d_model)
        linear(p + ".slf_attn.w_ks", n_head * d_k, d_model)
        linear(p + ".slf_attn.w_vs", n_head * d_v, d_model)
        ln(p + ".slf_attn.layer_norm")
        linear(p + ".slf_attn.fc", d_model, n_head * d_v)
        linear(p + ".pos_ffn.w_1", d_inner, d_model)
        linear(p + ".pos_ffn.w_2", d_model, d_inner)
        ln(p + ".pos_ffn.layer_norm")
    return sd


def synthetic_clips(n, t, seed=7, pad_frames=0):
    """LRW-shaped normalised gray clips [N,1,T,88,88] fp32: u8 ~ U{0..255} -> (u8/255 - 0.413621)/0.1700239
    (cvtransforms.py:44-48); the last `pad_frames` frames are all-zero like the reference's zero padding
    (data_gen.py:294-296)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    u8 = torch.randint(0, 256, (n, 1, t, 88, 88), generator=gen, dtype=torch.int32)
    x = (u8.float() / 255.0 - 0.413621) / 0.1700239
    if pad_frames:
        x[:, :, t - pad_frames:] = 0.0
    return x


def classifier_heads(seed=9, d_model=512, n_classes=1500, n_lang=2):
    """Synthetic weights of the stage-1 pre-training heads `fc_1500` / `fc_2`
    (VSR_visual_frontend_pretraining_on_LRW_LRW1000_classify/transformer/transformer.py:13-14), nn.Linear default
    init; used by the 1,000-clip top-1 parity check (BASELINE.json north_star)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    bound = 1.0 / math.sqrt(d_model)

    def u(*shape):
        return (torch.rand(shape, generator=gen) * 2.0 - 1.0) * bound

    return {"fc_1500.weight": u(n_classes, d_model), "fc_1500.bias": u(n_classes),
            "fc_2.weight": u(n_lang, d_model), "fc_2.bias": u(n_lang)}


def classify(enc_out, heads):
    """Word / language logits from encoder outputs [N,T,512], the evident intent of the stage-1 model's forward
    (…classify/transformer/transformer.py:31-36: pooled encoder output -> fc_1500, one frame -> fc_2): the word head
    reads the time average, the language head the last frame.  fp32 on whatever device `enc_out` lives on."""
    w1, b1 = heads["fc_1500.weight"].to(enc_out.device), heads["fc_1500.bias"].to(enc_out.device)
    w2, b2 = heads["fc_2.weight"].to(enc_out.device), heads["fc_2.bias"].to(enc_out.device)
    pooled = enc_out.float().mean(dim=1)
    return pooled @ w1.t() + b1, enc_out[:, -1].float() @ w2.t() + b2


def structured_clips(n, t, seed=11):
    """Synthetic clips with per-clip structure (iid-noise clips all map to nearly the same encoder output, which makes
    top-1 comparisons meaningless): each clip is a mixture of three drifting sinusoidal gratings with random
    orientation / frequency / speed / contrast, a brightness offset and a soft elliptical "mouth" blob whose opening
    oscillates over time, quantised to uint8 and normalised like the reference loader (cvtransforms.py:44-48).
    -> fp32 [n,1,t,88,88]."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 88), torch.linspace(-1, 1, 88), indexing="ij")
    tt = torch.arange(t, dtype=torch.float32).view(t, 1, 1)
    clips = []
    for _ in range(n):
        img = torch.zeros(t, 88, 88)
        for _g in range(3):
            th = torch.rand(1, generator=gen) * math.pi
            fr = 1.0 + 7.0 * torch.rand(1, generator=gen)
            sp = (torch.rand(1, generator=gen) - 0.5) * 1.2
            am = 0.15 + 0.35 * torch.rand(1, generator=gen)
            ph = torch.rand(1, generator=gen) * 2 * math.pi
            img = img + am * torch.sin(fr * math.pi * (xx * torch.cos(th) + yy * torch.sin(th)) + sp * tt + ph)
        cx, cy = (torch.rand(2, generator=gen) - 0.5) * 0.6
        ax = 0.25 + 0.35 * torch.rand(1, generator=gen)
        rate = 0.2 + 0.8 * torch.rand(1, generator=gen)
        ay = 0.08 + 0.25 * (0.5 + 0.5 * torch.sin(rate * tt + torch.rand(1, generator=gen) * 6.28))
        blob = torch.exp(-(((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2))
        img = img * (0.3 + 0.7 * torch.rand(1, generator=gen)) - (0.4 + 0.8 * torch.rand(1, generator=gen)) * blob
        img = img + (torch.rand(1, generator=gen) - 0.5) * 0.8 + 0.05 * torch.randn(t, 88, 88, generator=gen)
        u8 = ((img * 0.25 + 0.45).clamp(0, 1) * 255).round()
        clips.append(((u8 / 255.0 - 0.413621) / 0.1700239).unsqueeze(0))
    return torch.stack(clips).contiguous()


NORM_MEAN, NORM_STD = 0.413621, 0.1700239   # ColorNormalize, cvtransforms.py:44-48


def normalize_lut():
    """bf16 [256]: lut[u] = bf16(float32((u / 255. - mean) / std)), evaluated in float64 then cast to float32 exactly like
    the reference loader does (np.load(...) / 255. -> ColorNormalize -> float32 tensor, data_gen.py:122-125,276-296);
    used by the fused uint8 input path (ops.prep_clip_u8)."""
    import numpy as np
    u = np.arange(256, dtype=np.uint8)
    v = ((u / 255.) - NORM_MEAN) / NORM_STD          # float64, as numpy computes it in the reference
    return torch.from_numpy(v.astype(np.float32)).to(torch.bfloat16)


def synthetic_u8_clips(n, t, h0=96, w0=96, seed=21):
    """Raw loader-shaped clips: uint8 gray [n, t, h0, w0] (the reference's .npy files are [29, 96, 96] uint8)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randint(0, 256, (n, t, h0, w0), generator=gen, dtype=torch.int32).to(torch.uint8)
