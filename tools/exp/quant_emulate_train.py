#!/usr/bin/env python
"""CPU emulation of the TRAINING-mode rounding points of the frontend (raw conv output rounded to bf16, batch-stat BN,
output rounded to bf16) against the fp32 training forward, and of what that does to gradients (ReLU mask flips).
    python tools/exp/quant_emulate_train.py"""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sbl_for_multilingual_lip_reading_b200 import synth

class Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, on):
        return x.to(torch.bfloat16).float() if on else x
    @staticmethod
    def backward(ctx, g):
        return g, None            # gradient rounding is emulated separately (gq)

def rnd(x, on): return Round.apply(x, on)

def bn(x, sd, p, on):
    return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], training=True, eps=1e-5)

def forward(x, sd, q):
    w = sd["frontend3D.0.weight"]
    y = F.conv3d(rnd(x, q), rnd(w, q), stride=(1, 2, 2), padding=(2, 3, 3))
    y = rnd(F.relu(bn(rnd(y, q), sd, "frontend3D.1", q)), q)
    y = F.max_pool3d(y, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    y = y.transpose(1, 2).contiguous().view(-1, 64, 22, 22)
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for bi in range(2):
            p = f"resnet18.layer{li}.{bi}"
            st = stride if bi == 0 else 1
            res = y
            if bi == 0 and li != 1:
                res = rnd(bn(rnd(F.conv2d(y, rnd(sd[p + ".downsample.0.weight"], q), stride=st), q), sd, p + ".downsample.1", q), q)
            h = rnd(F.relu(bn(rnd(F.conv2d(y, rnd(sd[p + ".conv1.weight"], q), stride=st, padding=1), q), sd, p + ".bn1", q)), q)
            y = rnd(F.relu(bn(rnd(F.conv2d(h, rnd(sd[p + ".conv2.weight"], q), padding=1), q), sd, p + ".bn2", q) + res), q)
    return y.mean(dim=(2, 3))

def rel(a, b): return ((a - b).norm() / b.norm()).item()

torch.manual_seed(0)
sd = {k: v.clone().requires_grad_(v.dtype == torch.float32 and v.dim() > 0) for k, v in synth.frontend_state_dict(1).items()}
x = synth.structured_clips(4, 7, seed=31)
dy = torch.randn(28, 512, generator=torch.Generator().manual_seed(12))
outs, grads = {}, {}
for q in (False, True):
    for v in sd.values():
        if v.requires_grad: v.grad = None
    o = forward(x, sd, q)
    o.backward(dy)
    outs[q] = o.detach()
    grads[q] = {k: v.grad.clone() for k, v in sd.items() if v.requires_grad and v.grad is not None}
print("features: bf16-emulated training forward vs fp32 training forward:", rel(outs[True], outs[False]))
for k in ("resnet18.layer4.1.bn2.bias", "resnet18.layer4.1.conv2.weight", "resnet18.layer4.0.conv1.weight", "resnet18.layer3.1.conv1.weight",
          "resnet18.layer2.0.conv1.weight", "resnet18.layer1.0.conv1.weight", "frontend3D.0.weight"):
    print(f"  grad {k:40s} {rel(grads[True][k], grads[False][k]):.3f}  (fp32 backward formulas; only the FORWARD was rounded)")
