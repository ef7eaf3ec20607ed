#!/usr/bin/env python
"""Fused stem (sblk_stem_fused_fwd: clip prep inside the stem's producer warps) against prep_clip[_u8] + stem: bits, then
timings (L2 flushed between launches), full width and on a limited grid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops, synth

dev = torch.device("cuda")
ops.init()
sd = synth.frontend_state_dict(1)
w = sd["frontend3D.0.weight"].to(dev)
bn = [sd[f"frontend3D.1.{k}"].to(dev) for k in ("weight", "bias", "running_mean", "running_var")]
wp, bias = ops.pack_conv3d(w, *bn)
lut = synth.normalize_lut().to(dev)

ok = True
for (n, t) in ([] if os.environ.get("TIMING_ONLY") else [(1, 1), (1, 2), (1, 5), (2, 5), (3, 7), (2, 29), (1, 40), (32, 29)]):
    x = synth.synthetic_clips(n, t, seed=n * 100 + t).to(dev)
    for flat in (False, True):
        ref = ops.conv3d_bn_relu_pool(ops.prep_clip(x), wp, bias, flat=flat)
        got = ops.conv3d_bn_relu_pool(ops.raw_clip(x), wp, bias, flat=flat)
        a, b = (ref.data, got.data) if flat else (ref, got)
        same = torch.equal(a, b)
        ok &= same
        print(f"f32 N={n} T={t} flat={flat}: identical {same}", "" if same else f"diff elems {(a != b).sum().item()}")
    u8 = synth.synthetic_u8_clips(n, t, h0=100, w0=92, seed=n + t).to(dev)
    g = torch.Generator().manual_seed(n * 7 + t)
    offs = torch.stack([torch.randint(0, 13, (n * t,), generator=g), torch.randint(0, 5, (n * t,), generator=g)], 1).int().to(dev)
    for crop, tout in (((6, 2), t), ((0, 0), t + 3), (offs, t + 1)):
        ref = ops.conv3d_bn_relu_pool(ops.prep_clip_u8(u8, lut, tout, crop), wp, bias, flat=True)
        got = ops.conv3d_bn_relu_pool(ops.raw_clip_u8(u8, lut, tout, crop), wp, bias, flat=True)
        same = torch.equal(ref.data, got.data)
        ok &= same
        print(f"u8  N={n} T={t}->{tout} crop={'per-frame' if torch.is_tensor(crop) else crop}: identical {same}")
print("ALL IDENTICAL" if ok else "MISMATCH")

n, t = 32, 29
x = synth.synthetic_clips(n, t, seed=1).to(dev)
u8 = synth.synthetic_u8_clips(n, t, seed=2).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=12):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for lim in ((0,) if os.environ.get("TIMING_ONLY") else (0, 116, 84)):
    ops.set_sm_limit(lim)
    out = ops.conv3d_bn_relu_pool(ops.prep_clip(x), wp, bias, flat=True)
    xp = ops.prep_clip(x)
    xp8 = ops.prep_clip_u8(u8, lut, t, (4, 4))
    cases = {
        "prep f32": lambda: ops.prep_clip(x, out=xp[0]),
        "stem (prepped)": lambda: ops.conv3d_bn_relu_pool(xp, wp, bias, out=out.data, flat=True),
        "prep f32 + stem": lambda: ops.conv3d_bn_relu_pool(ops.prep_clip(x, out=xp[0]), wp, bias, out=out.data, flat=True),
        "fused f32": lambda: ops.conv3d_bn_relu_pool(ops.raw_clip(x), wp, bias, out=out.data, flat=True),
        "prep u8 + stem": lambda: ops.conv3d_bn_relu_pool(ops.prep_clip_u8(u8, lut, t, (4, 4), out=xp8[0]), wp, bias, out=out.data, flat=True),
        "fused u8": lambda: ops.conv3d_bn_relu_pool(ops.raw_clip_u8(u8, lut, t, (4, 4)), wp, bias, out=out.data, flat=True),
    }
    for k, fn in cases.items():
        fn()
        med, mn = timeit(fn)
        print(f"sm_limit {lim:3d}  {k:18s} median {med:7.1f} us  min {mn:7.1f} us")
ops.set_sm_limit(0)
