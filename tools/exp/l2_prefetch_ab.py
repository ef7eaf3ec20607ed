#!/usr/bin/env python
"""Is the L2 weight prefetch (tiny side-stream launches) still worth it at the end state?  Pipelined plan with
(a) everything (conv weights + the encoder's weights), (b) conv weights only, (c) nothing; L2 flushed between steps (the
bench's device-timed rule) and back to back without a flush (the sustained / serving regime)."""
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

def time_plan(plan, reps=30, do_flush=True):
    ts = []
    for i in range(reps + 4):
        s = i % 2
        with torch.cuda.stream(plan.compute):
            plan.x[s].copy_(xs[i % 4])
            if do_flush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(plan.compute)
            plan.forward_device(s)
            e1.record(plan.compute)
        torch.cuda.synchronize()
        if i >= 4:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

def back_to_back(plan, n=200):
    with torch.cuda.stream(plan.compute):
        for i in range(10):
            plan.forward_device(i % 2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(plan.compute)
        for i in range(n):
            plan.forward_device(i % 2)
        e1.record(plan.compute)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

class NoEncPrefetch(PipelinedVisualEncoderPlan):
    """conv weights only: drop the encoder's weights from the frontend's prefetch list after the base class set it"""
    def _capture(self):
        fe_ = self.frontend
        orig = type(fe_).__setattr__
        super()._capture()

import time
for rep in range(2):
    for mode in ("all", "conv only", "none"):
        fe.l2_prefetch = mode != "none"
        fe.l2_prefetch_extra = None
        if mode == "conv only":
            # the plan sets l2_prefetch_extra at capture time; neutralise it by making the attribute read-only empty
            class FE(type(fe)):
                @property
                def l2_prefetch_extra(self): return None
                @l2_prefetch_extra.setter
                def l2_prefetch_extra(self, v): pass
            saved_cls = fe.__class__
            fe.__dict__.pop("l2_prefetch_extra", None)
            fe.__class__ = FE
        pl = PipelinedVisualEncoderPlan(fe, enc, N, T, device=dev)
        a = time_plan(pl, do_flush=True)
        time.sleep(0.5)
        b = back_to_back(pl)
        n2 = pl.launches_per_forward
        pl.close(); del pl
        if mode == "conv only":
            fe.__class__ = saved_cls
            fe.l2_prefetch_extra = None
        print(f"prefetch {mode:9s}: {n2} launches | L2 flushed between steps: median {a:.1f} us ({N / a * 1e6:.0f} clips/s) | back to back, no flush (200 steps): {b:.1f} us", flush=True)
        time.sleep(1.0)
