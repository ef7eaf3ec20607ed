#!/usr/bin/env python
"""Per-stage clock64 stamps of the one-launch encoder stack (CTA 0 of cluster 0): where does the time go?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops, synth
from sbl_for_multilingual_lip_reading_b200.encoder import Encoder

dev = torch.device("cuda")
ops.init()
N, T, L = int(os.environ.get("N", 32)), int(os.environ.get("T", 29)), int(os.environ.get("L", 6))
enc = Encoder(512, L, 8, 64, 64, 512, 2048)
enc.load_state_dict(synth.encoder_state_dict(2, L))
enc = enc.to(dev).eval()
x16 = torch.randn(N * T, 512, device=dev).to(ops.enc16_dtype())
with torch.no_grad():
    pk = enc._get_packed()
    groups = (N + (128 // T) - 1) // (128 // T)
    dbg = torch.zeros(((1 + 4 * L) * 8 + 2 * groups,), dtype=torch.int64, device=dev)
    for it in range(3):
        ops.encoder_stack(x16, pk.stacked, N, T, debug_stamps=dbg, cluster_size=int(os.environ.get('CL', 0)),
                          groups_per_cluster=int(os.environ.get('GPC', 1)))
    torch.cuda.synchronize()
dall = dbg.cpu()
d = dall[:(1 + 4 * L) * 8].view(-1, 8)
t0 = int(d[0, 0])
mhz = float(os.environ.get("MHZ", 1965))
names = ["IN"] + [f"L{l}.{k}" for l in range(L) for k in ("QKV", "FC", "W1", "W2")]
print("stage     A-issue  loads-done  mma-issued  acc-ready  epi-done  stats-bar  ln-done  stage-end   (us since first A issue)")
for s in range(1 + 4 * L):
    r = [(int(v) - t0) / mhz if int(v) else float('nan') for v in d[s]]
    print(f"{names[s]:8s} {r[0]:8.2f} {r[5]:10.2f} {r[1]:11.2f} {r[2]:10.2f} {r[4]:9.2f} {r[6]:10.2f} {r[7]:8.2f} {r[3]:10.2f}")
cl = dall[(1 + 4 * L) * 8:].view(-1, 2)
base = int(cl[:, 0].min())
print("cluster start/end (us since first cluster start):", [(round((int(a) - base) / 1e3, 1), round((int(b) - base) / 1e3, 1)) for a, b in cl])
