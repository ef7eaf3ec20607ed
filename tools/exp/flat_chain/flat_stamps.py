#!/usr/bin/env python
"""Per-tile clock stamps of CTA 0 of the flat conv kernel (SBLK_DEBUG build, SBLK_FLAT_STAMPS=1): a single conv and,
when the library has it, the 4-conv chain.  python tools/exp/flat_stamps.py [C] [H]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 22
F = 928
dev = "cuda"
ops.init()
g = torch.Generator().manual_seed(0)
bf = torch.bfloat16
rows = ops.flat_rows(F, H, H)
x = ops.FlatActs(torch.randn(rows, C, generator=g).to(bf).to(dev), F, H, H)
ws = [ops.pack_flat_weight((torch.randn(C, 3, 3, C, generator=g) / (3 * C ** 0.5)).to(bf).to(dev)) for _ in range(4)]
bs = [torch.zeros(C, device=dev) for _ in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(2):
    ops.conv3x3_flat(x, ws[0], bs[0], relu=True)
torch.cuda.synchronize()
flush.zero_(); torch.cuda.synchronize()
os.environ["SBLK_FLAT_STAMPS"] = "1"
print("=== single conv", file=sys.stderr, flush=True)
ops.conv3x3_flat(x, ws[0], bs[0], relu=True)
torch.cuda.synchronize()
if hasattr(ops, "conv3x3_flat_chain") and os.environ.get("CHAIN"):
    del os.environ["SBLK_FLAT_STAMPS"]
    flags = ops.chain_flags(x, 4)
    lv = [(ws[0], bs[0], True, None), (ws[1], bs[1], True, "x"), (ws[2], bs[2], True, None), (ws[3], bs[3], True, 1)]
    n = 4 if C == 64 else 3
    if n == 3:
        lv = [(ws[0], bs[0], True, None), (ws[1], bs[1], True, None), (ws[2], bs[2], True, 0)]
    ops.conv3x3_flat_chain(x, lv[:n], flags)
    torch.cuda.synchronize()
    flush.zero_(); torch.cuda.synchronize()
    os.environ["SBLK_FLAT_STAMPS"] = "1"
    print("=== chain", file=sys.stderr, flush=True)
    ops.conv3x3_flat_chain(x, lv[:n], flags)
    torch.cuda.synchronize()
