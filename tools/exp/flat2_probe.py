#!/usr/bin/env python
"""flatconv2 timing probe (debug build): layer1 (C=64, 22x22) and layer2 (C=128, 11x11) shapes, normal MMAs against
N=32 MMAs (same issue stream, a fraction of the math and of the operand fetch), plus per-tile stamps of CTA 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops
DEV = "cuda"
ops.init()
g = torch.Generator().manual_seed(0)
bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

def timeit(fn, n=12):
    ts = []
    for i in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]

for (C, H) in ((64, 22), (128, 11)):
    F_ = 928
    rows = ops.flat_rows(F_, H, H)
    data = torch.randn(rows, C, generator=g).to(bf).to(DEV)
    xf = ops.FlatActs(data, F_, H, H)
    w = ops.pack_flat_weight((torch.randn(C, 3, 3, C, generator=g) / 24).to(bf).to(DEV))
    bias = torch.zeros(C, device=DEV)
    out = torch.empty_like(data)
    for mode in (0, 16):
        os.environ["SBLK_FLAT_DEBUG_MODE"] = str(mode)
        a = timeit(lambda: ops.conv3x3_flat(xf, w, bias, relu=True, out=out))
        b = timeit(lambda: ops.conv3x3_flat(xf, w, bias, relu=True, residual=xf, out=out))
        print(f"C={C} H={H} debug_mode={mode}: no-res {a:.1f} us, res {b:.1f} us", flush=True)
    os.environ["SBLK_FLAT_DEBUG_MODE"] = "0"
    os.environ["SBLK_FLAT_STAMPS"] = "1"
    ops.conv3x3_flat(xf, w, bias, relu=True, residual=xf, out=out)
    torch.cuda.synchronize()
    del os.environ["SBLK_FLAT_STAMPS"]
