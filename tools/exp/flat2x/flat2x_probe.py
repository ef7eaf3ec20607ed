#!/usr/bin/env python
"""Layer-1 flat conv (64 -> 64, 22 x 22, 928 frames): one output pixel per accumulator row (flatconv2_kernel<1>) against
two (flatconv2x_kernel).  Graph replay, PDL on, cold L2, chains of 1 / 2 / 4 / 8 convs -> marginal cost per conv; then
the plain and pipelined plans at the BASELINE shape with either variant.  python tools/exp/flat2x_probe.py"""
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
bf = torch.bfloat16
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
F_, H, C = 928, 22, 64


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


rows = ops.flat_rows(F_, H, H)
bufs = [ops.FlatActs(torch.randn(rows, C, generator=g).to(bf).to(dev), F_, H, H) for _ in range(3)]
w = ops.pack_flat_weight((torch.randn(C, 3, 3, C, generator=g) / (3 * C ** 0.5)).to(bf).to(dev))
bias = torch.zeros(C, device=dev)
flops = 2 * F_ * H * H * C * C * 9
for variant in (1, 0, 1, 0):
    ops.set_flat_variant(variant)
    for res in (False, True):
        def conv(i):
            ops.conv3x3_flat(bufs[i % 3], w, bias, relu=True, residual=bufs[(i + 2) % 3] if res else None,
                             out=bufs[(i + 1) % 3].data)
        r = [(n, graph_time(lambda: [conv(i) for i in range(n)])) for n in (1, 2, 4, 8)]
        per = (r[-1][1] - r[1][1]) / (r[-1][0] - r[1][0])
        print(f"variant {variant} ({'two' if variant == 0 else 'one'} px/row) residual={res}: " +
              "  ".join(f"{n}x {t:.1f} us" for n, t in r) +
              f"  -> marginal {per:.1f} us per conv ({flops / per / 1e6:.0f} TFLOP/s)", flush=True)

fe = visual_frontend(None); fe.load_state_dict(synth.frontend_state_dict(1))
enc = Encoder(512, 6, 8, 64, 64, 512, 2048); enc.load_state_dict(synth.encoder_state_dict(2, 6))
fe, enc = fe.to(dev).eval(), enc.to(dev).eval()
N, T = 32, 29
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
            out = plan.forward_device(s)
            e1.record(plan.compute)
        torch.cuda.synchronize()
        if i >= 4:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0], out.clone()


outs = {}
for rep in range(2):
    for variant in (1, 0):
        ops.set_flat_variant(variant)
        plain = VisualEncoderPlan(fe, enc, N, T, device=dev)
        med, best, _ = time_plan(plain)
        del plain
        pl = PipelinedVisualEncoderPlan(fe, enc, N, T, device=dev)
        med2, best2, o = time_plan(pl)
        outs[variant] = o
        pl.close()
        del pl
        print(f"flat variant {variant}: plain plan median {med:.1f} us best {best:.1f} | pipelined median {med2:.1f} us "
              f"best {best2:.1f} ({N / med2 * 1e6:.0f} clips/s)", flush=True)
print("pipelined outputs bit-identical across variants:", torch.equal(outs[0], outs[1]))
