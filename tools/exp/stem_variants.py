#!/usr/bin/env python
"""Stem A/B: the transposed stem (filter in tensor memory, variant 0) against the pixel-major stem (variant 1) and
against torch (Conv3d + BN + ReLU + MaxPool3d on the same bf16-rounded operands): outputs, then timings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from sbl_for_multilingual_lip_reading_b200 import ops, synth

dev = torch.device("cuda")
ops.init()
sd = synth.frontend_state_dict(1)
w = sd["frontend3D.0.weight"].to(dev)
bn = [sd[f"frontend3D.1.{k}"].to(dev) for k in ("weight", "bias", "running_mean", "running_var")]
wp, bias = ops.pack_conv3d(w, *bn)


def torch_ref(x):
    scale = bn[0] / torch.sqrt(bn[3] + 1e-5)
    wf = (w * scale.view(-1, 1, 1, 1, 1)).to(torch.bfloat16).float()
    y = F.conv3d(x.to(torch.bfloat16).float(), wf, None, (1, 2, 2), (2, 3, 3)) + (bn[1] - bn[2] * scale).view(1, -1, 1, 1, 1)
    y = F.max_pool3d(F.relu(y), (1, 3, 3), (1, 2, 2), (0, 1, 1))
    n, c, t, h, ww = y.shape
    return y.permute(0, 2, 3, 4, 1).reshape(n * t, h, ww, c)


ok = True
for (n, t) in [(1, 1), (1, 2), (1, 5), (2, 5), (3, 7), (2, 29), (1, 40), (32, 29)]:
    x = synth.synthetic_clips(n, t, seed=n * 100 + t).to(dev)
    xp = ops.prep_clip(x)
    outs = {}
    for var in (0, 1):
        ops.set_stem_variant(var)
        outs[var] = ops.conv3d_bn_relu_pool(xp, wp, bias).float()
        dirty = torch.full((ops.flat_rows(n * t, 22, 22), 64), 3.0, dtype=torch.bfloat16, device=dev)
        fl = ops.conv3d_bn_relu_pool(xp, wp, bias, out=dirty, flat=True)
        fd = fl.data.view(-1, 64).float()
        rows = ops.flat_rows(n * t, 22, 22)
        grid = fd.view(-1, 24, 64)[: n * t * 23 + 1]
        inner = grid[1:].view(n * t, 23, 24, 64)[:, :22, 1:23]
        halo_zero = (grid[0].abs().sum() + grid[1:].view(n * t, 23, 24, 64)[:, 22].abs().sum()
                     + grid[1:].view(n * t, 23, 24, 64)[:, :, 0].abs().sum()
                     + grid[1:].view(n * t, 23, 24, 64)[:, :, 23].abs().sum()).item()
        same_flat = torch.equal(inner, outs[var])
        if not same_flat or halo_zero != 0:
            ok = False
        print(f"N={n} T={t} variant {var}: flat == nhwc {same_flat}, halo sum {halo_zero}")
    ref = torch_ref(x)
    torch.cuda.synchronize()
    for var in (0, 1):
        err = ((outs[var] - ref).norm() / ref.norm()).item()
        mx = (outs[var] - ref).abs().max().item()
        print(f"   variant {var} vs torch: rel {err:.3e} max {mx:.3e}")
        if not err < 4e-3:
            ok = False
    d = (outs[0] - outs[1])
    print(f"   variant 0 vs 1: rel {(d.norm() / outs[1].norm()).item():.3e}, differing elements {(d != 0).float().mean().item():.3e}")
ops.set_stem_variant(0)

# timings (L2 flushed by a 256 MiB memset between launches)
n, t = 32, 29
x = synth.synthetic_clips(n, t, seed=1).to(dev)
xp = ops.prep_clip(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for lim in (0, 84):
    for var in (0, 1):
        ops.set_stem_variant(var)
        ops.set_sm_limit(lim)
        out = ops.conv3d_bn_relu_pool(xp, wp, bias, flat=True)
        ts = []
        for it in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv3d_bn_relu_pool(xp, wp, bias, out=out.data, flat=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts = sorted(ts[2:])
        us = ts[len(ts) // 2]
        print(f"N=32 T=29 sm_limit={lim} variant {var}: {us:.1f} us  ({2 * 64 * 44 * 44 * 245 * n * t / us / 1e6:.0f} TFLOP/s)")
ops.set_sm_limit(0)
ops.set_stem_variant(0)
print("ALL OK" if ok else "MISMATCH")
