#!/usr/bin/env python
"""End-state pipelined plan: cost of the placement gate and of the always-on dropout draw (F.dropout(ones) on the main
stream in front of the head)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops, synth
from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
from sbl_for_multilingual_lip_reading_b200.runner import PipelinedVisualEncoderPlan
from sbl_for_multilingual_lip_reading_b200.video_frontend import visual_frontend
dev = torch.device("cuda"); ops.init()
N, T = 32, 29
fe = visual_frontend(None); fe.load_state_dict(synth.frontend_state_dict(1))
enc = Encoder(512, 6, 8, 64, 64, 512, 2048); enc.load_state_dict(synth.encoder_state_dict(2, 6))
fe, enc = fe.to(dev).eval(), enc.to(dev).eval()
xs = [synth.synthetic_clips(N, T, seed=7 + i).to(dev) for i in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

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
    for gate in (True, False):
        for drop in (True, False):
            fe.always_on_dropout = drop
            pl = PipelinedVisualEncoderPlan(fe, enc, N, T, device=dev, gate=gate)
            med, best = time_plan(pl); n = pl.launches_per_forward; pl.close(); del pl
            print(f"gate={gate} always_on_dropout={drop}: {n} launches, median {med:.1f} us best {best:.1f} ({N / med * 1e6:.0f} clips/s)", flush=True)
fe.always_on_dropout = True
