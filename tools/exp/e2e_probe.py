#!/usr/bin/env python
"""Where does the end-to-end step (pinned fp32 clips in, host out) lose time against the device-timed step?
Times, inside a steady-state submit_host loop: the H2D copy of every step (events on the copy stream), the graph replay
(events on the compute stream) and the wall-clock step; then the same with the input copy split over two streams."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops, synth
from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
from sbl_for_multilingual_lip_reading_b200.runner import PipelinedVisualEncoderPlan
from sbl_for_multilingual_lip_reading_b200.video_frontend import visual_frontend

dev = torch.device("cuda")
ops.init()
N, T = 32, 29
fe = visual_frontend(None); fe.load_state_dict(synth.frontend_state_dict(1))
enc = Encoder(512, 6, 8, 64, 64, 512, 2048); enc.load_state_dict(synth.encoder_state_dict(2, 6))
fe, enc = fe.to(dev).eval(), enc.to(dev).eval()
pool = [synth.synthetic_clips(N, T, seed=7 + i).pin_memory() for i in range(4)]
out_host = [torch.empty((N, T, 512), dtype=torch.float32).pin_memory() for _ in range(2)]
plan = PipelinedVisualEncoderPlan(fe, enc, N, T, device=dev)
K = 60


def run(split):
    evs = []
    side2 = torch.cuda.Stream(device=dev)
    for i in range(K + 6):
        if i == 6:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        s = plan._i % 2
        plan._i += 1
        plan.copy_in.wait_event(plan.ev_done[s])
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x = pool[i % 4]
        if split:
            side2.wait_event(plan.ev_done[s])
            h = N // 2
            with torch.cuda.stream(plan.copy_in):
                a0.record(plan.copy_in)
                plan.x[s][:h].copy_(x[:h], non_blocking=True)
            with torch.cuda.stream(side2):
                plan.x[s][h:].copy_(x[h:], non_blocking=True)
                e2 = torch.cuda.Event(); e2.record(side2)
            with torch.cuda.stream(plan.copy_in):
                plan.copy_in.wait_event(e2)
                a1.record(plan.copy_in)
                plan.ev_in[s].record(plan.copy_in)
        else:
            with torch.cuda.stream(plan.copy_in):
                a0.record(plan.copy_in)
                plan.x[s].copy_(x, non_blocking=True)
                a1.record(plan.copy_in)
                plan.ev_in[s].record(plan.copy_in)
        plan.compute.wait_event(plan.ev_in[s])
        plan.compute.wait_event(plan.ev_out[s])
        with torch.cuda.stream(plan.compute):
            c0.record(plan.compute)
            plan.graphs[s].replay()
            c1.record(plan.compute)
            plan.ev_done[s].record(plan.compute)
        plan.copy_out.wait_event(plan.ev_done[s])
        with torch.cuda.stream(plan.copy_out):
            out_host[i % 2].copy_(plan.out[s ^ 1], non_blocking=True)
            plan.ev_out[s].record(plan.copy_out)
        if i >= 6:
            evs.append((a0, a1, c0, c1))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    h2d = sorted(a.elapsed_time(b) for a, b, _, _ in evs)
    rep = sorted(a.elapsed_time(b) for _, _, a, b in evs)
    print(f"split={split}: wall {1e3 * dt / K:.4f} ms/step ({N * K / dt:.0f} clips/s); H2D median {h2d[K // 2]:.4f} ms "
          f"(min {h2d[0]:.4f}, max {h2d[-1]:.4f}); replay median {rep[K // 2]:.4f} ms (min {rep[0]:.4f}, max {rep[-1]:.4f})")


for split in (False, True, False, True):
    run(split)
# device-only loop for reference (inputs already resident)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(plan.compute):
    for i in range(6):
        plan.forward_device(i % 2)
    e0.record(plan.compute)
    for i in range(K):
        plan.forward_device(i % 2)
    e1.record(plan.compute)
torch.cuda.synchronize()
print(f"device only, no flush: {e0.elapsed_time(e1) / K:.4f} ms/step")
