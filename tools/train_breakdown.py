#!/usr/bin/env python
"""Where one stage-1 training step (BASELINE configs[3] shard: 32 clips x 31 frames, fwd + bwd) spends its time:
per-libsblk-launch CUDA-event timing (ops.trace) aggregated by entry point, plus the wall/device total.
    python tools/train_breakdown.py [clips]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from sbl_for_multilingual_lip_reading_b200 import ops, stage1, synth
dev = torch.device("cuda:0"); ops.init()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = 31
torch.manual_seed(7)
m = stage1.Stage1Classifier(3, dropout=0.1).load_synthetic(1, 3).to(dev).train()
x = synth.synthetic_clips(n, T, seed=40, pad_frames=2).to(dev)
y = torch.randint(0, 1500, (n,)).to(dev); lang = torch.randint(0, 2, (n,)).to(dev)
def step():
    v_t, v_l = m(x)
    loss = F.cross_entropy(v_t, y) + 0.1 * F.cross_entropy(v_l, lang)
    m.zero_grad(); loss.backward()
    return loss
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record(); step(); e1.record(); torch.cuda.synchronize()
print(f"step (no trace): device {e0.elapsed_time(e1):.2f} ms, wall {1e3 * (time.perf_counter() - t0):.2f} ms")
sink = []
with ops.trace(sink):
    step()
torch.cuda.synchronize()
agg = {}
for r in sink:
    a = agg.setdefault(r["name"], [0, 0.0, 0.0])
    a[0] += 1; a[1] += r["start"].elapsed_time(r["end"]); a[2] += r["flops"]
tot = sum(a[1] for a in agg.values())
print(f"libsblk launches {len(sink)}, summed kernel time {tot:.2f} ms")
for k, (c, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:32s} x{c:4d}  {ms:8.3f} ms  {100 * ms / tot:5.1f}%  {fl / (ms * 1e-3) / 1e12 if fl else 0:7.1f} TFLOP/s")
# per-tag detail of the top entries
tags = {}
for r in sink:
    a = tags.setdefault((r["name"], r["tag"]), [0, 0.0])
    a[0] += 1; a[1] += r["start"].elapsed_time(r["end"])
print("top (entry, case):")
for (k, tg), (c, ms) in sorted(tags.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {k:28s} {tg:40s} x{c:3d} {ms:8.3f} ms")
