#!/usr/bin/env python
"""Two-stage pipeline probe (runner.PipelinedVisualEncoderPlan): L2-flushed replay time for a few (head SM limit, head
blocks, gate) settings next to the unpipelined plan (correctness: tests/test_parity_gpu.py).
python tools/exp/pipeline_probe.py [N] [T]"""
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


plain = VisualEncoderPlan(fe, enc, N, T, device=dev)
# ---- timing ----
med, best = time_plan(plain)
print(f"plain plan (split clusters {enc.split_clusters}): median {med:.1f} us  best {best:.1f} us")
variants = [(None, 0, True, None), (None, 0, False, None), (74, 0, True, None), (148, 0, True, None), (None, 1, True, None)]
if N * T <= 512:
    variants = [(None, hb, True, cl) for cl in (16, 8) for hb in (0, 2, 4, 8)] + [(148, 8, True, 16), (None, 8, False, 16)]
for limit, blocks, gate, cl in variants:
    try:
        p = PipelinedVisualEncoderPlan(fe, enc, N, T, device=dev, head_sm_limit=limit, head_blocks=blocks, gate=gate,
                                       enc_cluster=cl)
        med, best = time_plan(p)
        print(f"pipelined cl={p.enc_cluster} head_sm_limit={p.head_sm_limit} head_blocks={blocks} gate={gate}: median "
              f"{med:.1f} us  best {best:.1f} us  ({N / med * 1e6:.0f} clips/s)")
        del p
    except Exception as e:  # noqa: BLE001
        print(f"pipelined cl={cl} limit={limit} blocks={blocks} gate={gate}: FAILED {e}")
