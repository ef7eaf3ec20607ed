#!/usr/bin/env python
"""Layer-3 / layer-4 residual blocks as one launch (ops.conv_block) against the two launches: block marginal cost in a graph, then
the plain / pipelined plans with Lipreading.fuse_blocks on / off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops, synth
from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
from sbl_for_multilingual_lip_reading_b200.runner import PipelinedVisualEncoderPlan, VisualEncoderPlan
from sbl_for_multilingual_lip_reading_b200.video_frontend import visual_frontend
dev = torch.device("cuda"); ops.init()
N, T = 32, 29
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

for (cin, cout, h, stride) in ((128, 256, 11, 2), (256, 256, 6, 1), (256, 512, 6, 2), (512, 512, 3, 1)):
    x = torch.randn(F_, h, h, cin, generator=g).to(bf).to(dev)
    w1 = (torch.randn(cout, 3, 3, cin, generator=g) / (9 * cin) ** 0.5).to(bf).to(dev)
    w2 = (torch.randn(cout, 3, 3, cout, generator=g) / (9 * cout) ** 0.5).to(bf).to(dev)
    wd = (torch.randn(cout, 1, 1, cin, generator=g) / cin ** 0.5).to(bf).to(dev) if stride == 2 else None
    b = torch.zeros(cout, device=dev)
    p = (h - 1) // stride + 1
    outs = [torch.empty(F_, p, p, cout, dtype=bf, device=dev) for _ in range(2)]
    y1 = torch.empty(F_, p, p, cout, dtype=bf, device=dev)
    def two(i):
        y = ops.conv2d(x, w1, b, stride=stride, relu=True)
        if wd is not None:
            ops.conv2d(y, w2, b, stride=1, relu=True, ext=(x, wd, stride), out=outs[i % 2])
        else:
            ops.conv2d(y, w2, b, stride=1, relu=True, residual=x, out=outs[i % 2])
    def one(i):
        ops.conv_block(x, w1, b, w2, b, w_ds=wd, stride=stride, out=outs[i % 2], y1=y1)
    for name, fn in (("two launches", two), ("one launch", one)):
        res = [(n, graph_time(lambda: [fn(i) for i in range(n)])) for n in (1, 2, 4)]
        per = (res[-1][1] - res[0][1]) / (res[-1][0] - res[0][0])
        print(f"{cin}->{cout} H={h} s{stride}  {name}: " + "  ".join(f"{n}x {t:.1f} us" for n, t in res) + f"  -> marginal {per:.1f} us", flush=True)
    two(0); one(1); torch.cuda.synchronize()
    print("   identical:", torch.equal(outs[0], outs[1]), flush=True)

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
    for fuse in (0, 256, 512):
        fe.fuse_blocks, fe.fuse_blocks_max = bool(fuse), fuse or 512
        plain = VisualEncoderPlan(fe, enc, N, T, device=dev)
        med, best = time_plan(plain); del plain
        pl = PipelinedVisualEncoderPlan(fe, enc, N, T, device=dev)
        med2, best2 = time_plan(pl); pl.close(); del pl
        print(f"fuse_blocks up to {fuse}: plain plan median {med:.1f} best {best:.1f} | pipelined median {med2:.1f} best {best2:.1f} ({N / med2 * 1e6:.0f} clips/s)", flush=True)
