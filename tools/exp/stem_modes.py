#!/usr/bin/env python
"""Timing experiments on the transposed stem (needs a -DSBLK_DEBUG build): SBLK_C3D_DEBUG_MODE = 0..4."""
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
n, t = 32, 29
x = synth.synthetic_clips(n, t, seed=1).to(dev)
xp = ops.prep_clip(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for var in (0,):
    ops.set_stem_variant(var)
    for mode in [int(v) for v in os.environ.get('MODES', '0,1,2,4,6,8,16,32').split(',')]:
        os.environ["SBLK_C3D_DEBUG_MODE"] = str(mode)
        out = ops.conv3d_bn_relu_pool(xp, wp, bias, flat=True)
        ts = []
        for it in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv3d_bn_relu_pool(xp, wp, bias, out=out.data, flat=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts = sorted(ts[2:])
        print(f"variant {var} debug_mode {mode}: {ts[len(ts) // 2]:.1f} us")

# per-CTA clock stamps (debug build): where does a CTA's time go?
os.environ["SBLK_C3D_DEBUG_MODE"] = os.environ.get("STAMP_MODE", "0")
dbg = torch.zeros((148, 8), dtype=torch.int64, device=dev)
os.environ["SBLK_C3D_STAMPS"] = hex(dbg.data_ptr())
for it in range(3):
    flush.zero_()
    ops.conv3d_bn_relu_pool(xp, wp, bias, out=out.data, flat=True)
torch.cuda.synchronize()
d = dbg.cpu().double()
names = ["entry", "at grid wait", "after grid wait", "first data", "MMA issued", "epilogue done", "loader done"]
rel = (d[:, 1:7] - d[:, :1]) / 1965.0
for k in range(6):
    print(f"  {names[k + 1]:16s}: mean {rel[:, k].mean():7.2f} us  min {rel[:, k].min():7.2f}  max {rel[:, k].max():7.2f}  (since CTA entry)")
del os.environ["SBLK_C3D_STAMPS"]
