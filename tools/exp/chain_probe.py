#!/usr/bin/env python
"""How much do kernel boundaries cost inside a graph?  Chains of identical stride-1 convs (layer1 ... layer4 shapes) are
captured into one CUDA graph with PDL on, replayed with cold L2, and compared with the steady-state MMA time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops
DEV = "cuda"
ops.init()
g = torch.Generator().manual_seed(0)
bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
F_ = 928


def graph_time(fn, reps=12):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
        s.synchronize()
        gr = torch.cuda.CUDAGraph()
        old = ops.set_pdl(True)
        with torch.cuda.graph(gr, stream=s):
            fn()
        ops.set_pdl(old)
        ts = []
        for i in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); gr.replay(); e1.record(s)
            s.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]


SHAPES = [tuple(int(v) for v in t.split(":")) for t in os.environ.get("SHAPES", "64:22:1,128:11:1,256:6:0,512:3:0").split(",")]
for (C, H, flat) in SHAPES:
    if flat:
        rows = ops.flat_rows(F_, H, H)
        bufs = [ops.FlatActs(torch.randn(rows, C, generator=g).to(bf).to(DEV), F_, H, H) for _ in range(3)]
        w = ops.pack_flat_weight((torch.randn(C, 3, 3, C, generator=g) / (3 * C ** 0.5)).to(bf).to(DEV))
        bias = torch.zeros(C, device=DEV)
        def conv(i):
            ops.conv3x3_flat(bufs[i % 3], w, bias, relu=True, out=bufs[(i + 1) % 3].data)
    else:
        bufs = [torch.randn(F_, H, H, C, generator=g).to(bf).to(DEV) for _ in range(3)]
        w = (torch.randn(C, 3, 3, C, generator=g) / (3 * C ** 0.5)).to(bf).to(DEV)
        bias = torch.zeros(C, device=DEV)
        def conv(i):
            ops.conv2d(bufs[i % 3], w, bias, stride=1, relu=True, out=bufs[(i + 1) % 3])
    flops = 2 * F_ * H * H * C * C * 9
    res = []
    for n in (1, 2, 4, 8):
        t = graph_time(lambda: [conv(i) for i in range(n)])
        res.append((n, t))
    per = (res[-1][1] - res[1][1]) / (res[-1][0] - res[1][0])
    print(f"C={C} H={H}: " + "  ".join(f"{n} convs {t:.1f} us" for n, t in res) +
          f"  -> marginal {per:.1f} us per conv ({flops / per / 1e6:.0f} TFLOP/s), first conv {res[0][1]:.1f} us", flush=True)

# strided block heads (3x3/s2 conv + 1x1/s2 downsample in one launch)
for (Cin, Cout, H) in (() if os.environ.get("NO_HEADS") else ((64, 128, 22), (128, 256, 11), (256, 512, 6))):
    x = torch.randn(F_, H, H, Cin, generator=g).to(bf).to(DEV)
    w = (torch.randn(Cout, 3, 3, Cin, generator=g) / (3 * Cin ** 0.5)).to(bf).to(DEV)
    wd = (torch.randn(Cout, 1, 1, Cin, generator=g) / (Cin ** 0.5)).to(bf).to(DEV)
    b1 = torch.zeros(Cout, device=DEV); b2 = torch.zeros(Cout, device=DEV)
    def head(i):
        ops.conv2d_dual(x, w, b1, wd, b2, stride=2, relu=True)
    flops = 2 * F_ * (H // 2) * (H // 2) * Cout * Cin * 10
    res = []
    for n in (1, 2, 4, 8):
        res.append((n, graph_time(lambda: [head(i) for i in range(n)])))
    per = (res[-1][1] - res[1][1]) / (res[-1][0] - res[1][0])
    print(f"head {Cin}->{Cout} H={H}: " + "  ".join(f"{n}x {t:.1f} us" for n, t in res) +
          f"  -> marginal {per:.1f} us ({flops / per / 1e6:.0f} TFLOP/s)", flush=True)
    if os.environ.get("SBLK_IGEMM2_STAMPS_RUN"):
        os.environ["SBLK_IGEMM2_STAMPS"] = "1"
        head(0); torch.cuda.synchronize()
        del os.environ["SBLK_IGEMM2_STAMPS"]
