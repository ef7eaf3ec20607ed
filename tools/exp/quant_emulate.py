#!/usr/bin/env python
"""CPU emulation of the operand rounding points of the libsblk path (no GPU needed): where does the relative error of
the encoder output come from, and what would 16-bit operand formats with more mantissa (fp16) buy?

Every contraction is evaluated in fp32 on operands rounded to the emulated format at exactly the places the kernels
round (packed weights incl. the BN fold, stored activations, q/k/v tiles, softmax probabilities, attention output,
FFN hidden); LayerNorm / softmax / residual stream stay fp32 like the kernels.  Compared against the fp32 oracle.

    python tools/exp/quant_emulate.py [clips]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import visual_encoder_oracle as O  # noqa: E402
from sbl_for_multilingual_lip_reading_b200 import synth  # noqa: E402


def rnd(x, dt):
    return x if dt is None else x.to(dt).float()


def fold(w, sd, p, dt):
    s = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + 1e-5)
    b = sd[p + ".bias"] - sd[p + ".running_mean"] * s
    return rnd(w * s.view(-1, *([1] * (w.dim() - 1))), dt), b


def frontend(x, sd, dt, dt_in):
    w, b = fold(sd["frontend3D.0.weight"], sd, "frontend3D.1", dt)
    y = F.conv3d(rnd(x, dt_in), w, b, stride=(1, 2, 2), padding=(2, 3, 3))
    y = rnd(F.relu(y), dt)          # the kernel rounds before the max-pool (max commutes with rounding)
    y = F.max_pool3d(y, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    y = y.transpose(1, 2).contiguous().view(-1, 64, 22, 22)
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for bi in range(2):
            p = f"resnet18.layer{li}.{bi}"
            st = stride if bi == 0 else 1
            w1, b1 = fold(sd[p + ".conv1.weight"], sd, p + ".bn1", dt)
            w2, b2 = fold(sd[p + ".conv2.weight"], sd, p + ".bn2", dt)
            res = y
            if bi == 0 and li != 1:
                wd, bd = fold(sd[p + ".downsample.0.weight"], sd, p + ".downsample.1", dt)
                res = rnd(F.conv2d(y, wd, bd, stride=st), dt)
            h = rnd(F.relu(F.conv2d(y, w1, b1, stride=st, padding=1)), dt)
            y = rnd(F.relu(F.conv2d(h, w2, b2, padding=1) + res), dt)
    return y.mean(dim=(2, 3))


def encoder(feat, sd, dt, n_layers=6, h=8, dk=64):
    n, t, _ = feat.shape
    lin = lambda x, p: F.linear(x, rnd(sd[p + ".weight"], dt), sd[p + ".bias"])  # noqa: E731
    ln = lambda x, p: F.layer_norm(x, (512,), sd[p + ".weight"], sd[p + ".bias"], 1e-5)  # noqa: E731
    x = ln(lin(rnd(feat, dt), "linear_in"), "layer_norm_in") + sd["positional_encoding.pe"][:, :t]
    for i in range(n_layers):
        a, f = f"layer_stack.{i}.slf_attn", f"layer_stack.{i}.pos_ffn"
        x16 = rnd(x, dt)
        q = rnd(lin(x16, a + ".w_qs"), dt).view(n, t, h, dk).permute(0, 2, 1, 3)
        k = rnd(lin(x16, a + ".w_ks"), dt).view(n, t, h, dk).permute(0, 2, 1, 3)
        v = rnd(lin(x16, a + ".w_vs"), dt).view(n, t, h, dk).permute(0, 2, 1, 3)
        pr = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
        att = rnd((rnd(pr, dt) @ v).permute(0, 2, 1, 3).reshape(n, t, 512), dt)
        x = ln(lin(att, a + ".fc") + x, a + ".layer_norm")
        hid = rnd(F.relu(lin(rnd(x, dt), f + ".w_1")), dt)
        x = ln(lin(hid, f + ".w_2") + x, f + ".layer_norm")
    return x


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    t = 29
    torch.set_grad_enabled(False)
    fsd, esd = synth.frontend_state_dict(1), synth.encoder_state_dict(2, 6)
    x = synth.structured_clips(n, t, seed=5000)
    ref_feat = O.frontend_forward(x, fsd).view(n, t, 512)
    ref_out = O.encoder_forward(ref_feat, [t] * n, esd)[0]
    bf, hf = torch.bfloat16, torch.float16
    print(f"{n} clips x {t} frames; relative Frobenius error vs the fp32 oracle")
    for name, fdt, fin, edt in (("bf16 frontend + bf16 encoder (round 1)", bf, bf, bf),
                                ("bf16 frontend + fp16 encoder", bf, bf, hf),
                                ("fp16 frontend (bf16 clip) + fp16 encoder", hf, bf, hf),
                                ("fp16 frontend + fp16 encoder", hf, hf, hf),
                                ("fp32 frontend + bf16 encoder", None, None, bf),
                                ("fp32 frontend + fp16 encoder", None, None, hf)):
        feat = frontend(x, fsd, fdt, fin).view(n, t, 512)
        out = encoder(feat, esd, edt)
        print(f"  {name:45s} features {rel(feat, ref_feat):.2e}   encoder output {rel(out, ref_out):.2e}")


if __name__ == "__main__":
    main()
