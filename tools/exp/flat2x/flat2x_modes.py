#!/usr/bin/env python
"""SBLK_DEBUG build only: flatconv2x_kernel timing switches and per-tile stamps (layer-1 shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops
dev = torch.device("cuda"); ops.init()
bf = torch.bfloat16; g = torch.Generator().manual_seed(0)
F_, H, C = 928, 22, 64
rows = ops.flat_rows(F_, H, H)
bufs = [ops.FlatActs(torch.randn(rows, C, generator=g).to(bf).to(dev), F_, H, H) for _ in range(3)]
w = ops.pack_flat_weight((torch.randn(C, 3, 3, C, generator=g) / (3 * C ** 0.5)).to(bf).to(dev))
bias = torch.zeros(C, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def t(res, reps=12):
    ts = []
    for i in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv3x3_flat(bufs[0], w, bias, relu=True, residual=bufs[2] if res else None, out=bufs[1].data)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]

for variant in (1, 0):
    ops.set_flat_variant(variant)
    for mode in ([0] if variant == 1 else [0, 1, 2, 3, 4]):
        os.environ["SBLK_FLAT_DEBUG_MODE"] = str(mode)
        print(f"variant {variant} mode {mode}: no residual {t(False):.1f} us, residual {t(True):.1f} us", flush=True)
    os.environ["SBLK_FLAT_DEBUG_MODE"] = "0"
ops.set_flat_variant(0)
for res in (False, True):
    os.environ["SBLK_FLAT_STAMPS"] = "1"
    print("stamps, residual =", res, flush=True)
    ops.conv3x3_flat(bufs[0], w, bias, relu=True, residual=bufs[2] if res else None, out=bufs[1].data)
    torch.cuda.synchronize()
    del os.environ["SBLK_FLAT_STAMPS"]
