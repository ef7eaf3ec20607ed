#!/usr/bin/env python
"""Alternating tile direction of the flat convs (Lipreading.alternate_tile_order): frontend alone (graph, PDL on, cold
L2), plain plan, pipelined plan, both settings, twice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops, synth
from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
from sbl_for_multilingual_lip_reading_b200.runner import PipelinedVisualEncoderPlan, VisualEncoderPlan
from sbl_for_multilingual_lip_reading_b200.video_frontend import visual_frontend
dev = torch.device("cuda"); ops.init()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 29
fe = visual_frontend(None); fe.load_state_dict(synth.frontend_state_dict(1))
enc = Encoder(512, 6, 8, 64, 64, 512, 2048); enc.load_state_dict(synth.encoder_state_dict(2, 6))
fe, enc = fe.to(dev).eval(), enc.to(dev).eval()
xs = [synth.synthetic_clips(N, T, seed=7 + i).to(dev) for i in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def graph_time(fn, reps=16):
    s = torch.cuda.Stream()
    with torch.no_grad(), torch.cuda.stream(s):
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
    for alt in (False, True):
        fe.alternate_tile_order = alt
        t_fe = graph_time(lambda: fe._frontend_forward(xs[0]))
        plain = VisualEncoderPlan(fe, enc, N, T, device=dev)
        med, best = time_plan(plain)
        del plain
        pl = PipelinedVisualEncoderPlan(fe, enc, N, T, device=dev)
        med2, best2 = time_plan(pl)
        pl.close(); del pl
        print(f"alternate_tile_order={alt}: frontend {t_fe:.1f} us | plain plan median {med:.1f} best {best:.1f} | "
              f"pipelined median {med2:.1f} best {best2:.1f} ({N / med2 * 1e6:.0f} clips/s)", flush=True)
