#!/usr/bin/env python
"""Timing experiments: flatconv / stem / igemm with the debug modes (wrong results, timing only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops
DEV = "cuda"
ops.init()
g = torch.Generator().manual_seed(0)
bf = torch.bfloat16

def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

F_, H = 928, 22
rows = ops.flat_rows(F_, H, H)
data = torch.randn(rows, 64, generator=g).to(bf).to(DEV)
xf = ops.FlatActs(data, F_, H, H)
w = ops.pack_flat_weight((torch.randn(64, 3, 3, 64, generator=g) / 24).to(bf).to(DEV))
bias = torch.zeros(64, device=DEV)
out = torch.empty_like(data)
for pdl in (0, 1):
    ops.set_pdl(bool(pdl))
    for mode in (0, 1, 2, 3):
        os.environ["SBLK_FLAT_DEBUG_MODE"] = str(mode)
        a = timeit(lambda: ops.conv3x3_flat(xf, w, bias, relu=True, out=out))
        b = timeit(lambda: ops.conv3x3_flat(xf, w, bias, relu=True, residual=xf, out=out))
        print(f"flat pdl={pdl} mode={mode} (1=no A loads, 2=no epilogue): no-res {a:.1f} us, res {b:.1f} us", flush=True)
os.environ["SBLK_FLAT_DEBUG_MODE"] = "0"
