#!/usr/bin/env python
"""Pipelined plan: L2-flushed replay time against head_frac (share of layer1.0.conv1's frames that runs inside the head)
and the stem variant.  python tools/exp/head_frac_probe.py [N] [T]"""
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


fe.fuse_prep = bool(int(os.environ.get("FUSE_PREP", "1")))
print("fuse_prep", fe.fuse_prep)
for var in [int(v) for v in os.environ.get('VARIANTS', '0,1').split(',')]:
    ops.set_stem_variant(var)
    plain = VisualEncoderPlan(fe, enc, N, T, device=dev)
    med, best = time_plan(plain)
    print(f"stem variant {var}: plain plan median {med:.1f} us  best {best:.1f} us")
    del plain
    for hf in [(float(v) if v != "auto" else None) for v in os.environ.get("HEAD_FRACS", "0.5,0.7,0.85,1.0").split(",")]:
        for lim in [int(v) for v in os.environ.get("HEAD_LIMITS", "0").split(",")]:
            p = PipelinedVisualEncoderPlan(fe, enc, N, T, device=dev, head_frac=hf, head_sm_limit=lim or None,
                                           head_blocks=(int(os.environ["HEAD_BLOCKS"]) if "HEAD_BLOCKS" in os.environ else None),
                                           enc_gpc=(int(os.environ["ENC_GPC"]) if "ENC_GPC" in os.environ else None))
            med, best = time_plan(p)
            print(f"stem variant {var}: pipelined head_frac={p.head_frac} head_blocks={p.head_blocks} gpc={p.enc_gpc} head_sm_limit={p.head_sm_limit}: median {med:.1f} us  best {best:.1f} us  "
                  f"({N / med * 1e6:.0f} clips/s)")
            p.close()
            del p
ops.set_stem_variant(0)
