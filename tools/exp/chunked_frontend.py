#!/usr/bin/env python
"""Does depth-first (clip-group) scheduling of the frontend help L2 residency? Graph-timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops, synth
from sbl_for_multilingual_lip_reading_b200.video_frontend import visual_frontend
dev = torch.device("cuda")
ops.init(); ops.set_pdl(True)
N, T = 32, 29
fe = visual_frontend(None); fe.load_state_dict(synth.frontend_state_dict(1)); fe = fe.to(dev).eval()
fe.always_on_dropout = False
x = synth.synthetic_clips(N, T, seed=7).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def graph_time(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.no_grad(), torch.cuda.stream(s):
        for _ in range(3):
            fn()
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g, stream=s):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3

for G in (1, 2, 4, 8):
    b = [(g * N) // G for g in range(G + 1)]
    print(f"frontend in {G} sequential clip groups: {graph_time(lambda: [fe._frontend_forward(x[b[g]:b[g+1]]) for g in range(G)]):.1f} us", flush=True)
