#!/usr/bin/env python
"""Block heads of layers 3-4: the dual kernel (conv1 + downsample on 128-wide pair tiles, branch as bf16 residual of
conv2) against the folded form (conv1 alone on 256-wide pair tiles, branch as a K-extension of conv2).
1. whole residual block in a graph (PDL on, cold L2), chains of 1 / 2 / 4 blocks -> marginal cost per block;
2. plain and pipelined plans at the BASELINE shape with Lipreading.fold_downsample on / off.
python tools/exp/fold_ds_probe.py [N] [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops, synth
from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
from sbl_for_multilingual_lip_reading_b200.runner import PipelinedVisualEncoderPlan, VisualEncoderPlan
from sbl_for_multilingual_lip_reading_b200.video_frontend import visual_frontend

dev = torch.device("cuda")
ops.init()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 29
F_ = N * T
bf = torch.bfloat16
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def graph_time(fn, reps=14):
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


for (cin, cout, h) in ((128, 256, 11), (256, 512, 6)):
    p = (h - 1) // 2 + 1
    x = torch.randn(F_, h, h, cin, generator=g).to(bf).to(dev)
    w1 = (torch.randn(cout, 3, 3, cin, generator=g) / (3 * cin ** 0.5)).to(bf).to(dev)
    w2 = (torch.randn(cout, 3, 3, cout, generator=g) / (3 * cout ** 0.5)).to(bf).to(dev)
    wd = (torch.randn(cout, 1, 1, cin, generator=g) / cin ** 0.5).to(bf).to(dev)
    b = torch.zeros(cout, device=dev)
    outs = [torch.empty(F_, p, p, cout, dtype=bf, device=dev) for _ in range(2)]

    def dual(i):
        y, res = ops.conv2d_dual(x, w1, b, wd, b, stride=2, relu=True)
        ops.conv2d(y, w2, b, stride=1, relu=True, residual=res, out=outs[i % 2])

    def folded(i):
        y = ops.conv2d(x, w1, b, stride=2, relu=True)
        ops.conv2d(y, w2, b, stride=1, relu=True, ext=(x, wd, 2), out=outs[i % 2])

    def conv1_only(i):
        ops.conv2d(x, w1, b, stride=2, relu=True)

    def dual_only(i):
        ops.conv2d_dual(x, w1, b, wd, b, stride=2, relu=True)

    flops = 2 * F_ * p * p * cout * (10 * cin + 9 * cout)
    for name, fn in (("dual head + conv2", dual), ("conv1 + conv2 with folded branch", folded),
                     ("dual head alone", dual_only), ("conv1 (256-wide tiles) alone", conv1_only)):
        res = [(n, graph_time(lambda: [fn(i) for i in range(n)])) for n in (1, 2, 4)]
        per = (res[-1][1] - res[0][1]) / (res[-1][0] - res[0][0])
        print(f"{cin}->{cout} H={h}  {name}: " + "  ".join(f"{n}x {t:.1f} us" for n, t in res) +
              f"  -> marginal {per:.1f} us", flush=True)
    dual(0); folded(1); torch.cuda.synchronize()
    d = (outs[0].float() - outs[1].float()).norm() / outs[0].float().norm()
    print(f"   folded vs dual block output: rel diff {d.item():.2e}", flush=True)

fe = visual_frontend(None); fe.load_state_dict(synth.frontend_state_dict(1))
enc = Encoder(512, 6, 8, 64, 64, 512, 2048); enc.load_state_dict(synth.encoder_state_dict(2, 6))
fe, enc = fe.to(dev).eval(), enc.to(dev).eval()
xs = [synth.synthetic_clips(N, T, seed=7 + i).to(dev) for i in range(4)]


def time_plan(plan, reps=30):
    ts = []
    for i in range(reps + 4):
        s = i % 2
        with torch.cuda.stream(plan.compute):
            plan.x[s].copy_(xs[i % 4])
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(plan.compute)
            plan.forward_device(s)
            e1.record(plan.compute)
        torch.cuda.synchronize()
        if i >= 4:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for rep in range(2):
    for fold in (False, True):
        fe.fold_downsample = fold
        plain = VisualEncoderPlan(fe, enc, N, T, device=dev)
        med, best = time_plan(plain)
        del plain
        pl = PipelinedVisualEncoderPlan(fe, enc, N, T, device=dev)
        med2, best2 = time_plan(pl)
        pl.close()
        del pl
        print(f"fold_downsample={fold}: plain plan median {med:.1f} us best {best:.1f} | pipelined median {med2:.1f} us "
              f"best {best2:.1f} ({N / med2 * 1e6:.0f} clips/s)", flush=True)
