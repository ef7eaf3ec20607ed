#!/usr/bin/env python
"""What would a third pipeline stage buy?  Pipelined plan whose graphs contain NO clip prep (the stem reads a prepped
clip made outside the graph) — (A) nothing else, i.e. the upper bound; (B) with the prep of the NEXT batch launched on a
side stream inside the same timed step (prep(i+1) || frontend(i) || encoder(i-1)); against the product plan (fused prep)."""
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
side = torch.cuda.Stream()

def time_plan(plan, prep_bufs=None, reps=30, when="start"):
    ts = []
    for i in range(reps + 4):
        s = i % 2
        with torch.cuda.stream(plan.compute):
            plan.x[s].copy_(xs[i % 4])
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(plan.compute)
            if prep_bufs is not None and when == "start":
                side.wait_event(e0)
                with torch.cuda.stream(side):
                    ops.prep_clip(xs[(i + 1) % 4], out=prep_bufs[s ^ 1])
            plan.forward_device(s)
            if prep_bufs is not None and when == "end":
                side.wait_event(e0)
                with torch.cuda.stream(side):
                    ops.prep_clip(xs[(i + 1) % 4], out=prep_bufs[s ^ 1])
            if prep_bufs is not None:
                plan.compute.wait_stream(side)
            e1.record(plan.compute)
        torch.cuda.synchronize()
        if i >= 4:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

for rep in range(2):
    pl = PipelinedVisualEncoderPlan(fe, enc, N, T, device=dev)
    med, best = time_plan(pl); pl.close(); del pl
    print(f"product plan (prep fused into the stem): median {med:.1f} us best {best:.1f} ({N / med * 1e6:.0f} clips/s)", flush=True)
    bufs = [ops.prep_clip(xs[k])[0] for k in range(2)]
    orig = fe._frontend_chain
    cur = {"s": 0}
    def chain(x, pk, feat_out, chain_id, prep=None):
        s = 0 if x.data_ptr() == PL_X[0] else 1
        return orig(x, pk, feat_out, chain_id, prep=lambda xx: (bufs[s], N, T))
    fe.fuse_prep = False
    fe._frontend_chain = chain
    PL_X = [0, 0]
    # the plan allocates x[] in its constructor and captures immediately: resolve the slot by pointer lazily
    class P(PipelinedVisualEncoderPlan):
        def _capture(self):
            PL_X[0], PL_X[1] = self.x[0].data_ptr(), self.x[1].data_ptr()
            super()._capture()
    pl = P(fe, enc, N, T, device=dev)
    med, best = time_plan(pl)
    print(f"(A) no prep anywhere (stem reads a prepped clip): median {med:.1f} us best {best:.1f} ({N / med * 1e6:.0f} clips/s)", flush=True)
    med, best = time_plan(pl, bufs, when="start")
    print(f"(B) + prep of the next batch on a side stream, issued in front of the replay: median {med:.1f} us best {best:.1f} ({N / med * 1e6:.0f} clips/s)", flush=True)
    med, best = time_plan(pl, bufs, when="end")
    print(f"(B') + prep of the next batch on a side stream, issued behind the replay: median {med:.1f} us best {best:.1f} ({N / med * 1e6:.0f} clips/s)", flush=True)
    pl.close(); del pl
    fe._frontend_chain = orig
    fe.fuse_prep = True
