#!/usr/bin/env python
"""Timing of one greedy decode of the libsblk SBL decoder under its execution options.
    python tools/decoder_breakdown.py [clips] [frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops
from sbl_for_multilingual_lip_reading_b200.decoder import Decoder
dev = torch.device("cuda:0"); ops.init()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
t = int(sys.argv[2]) if len(sys.argv) > 2 else 30
torch.manual_seed(7)
dec = Decoder(0, 1, 58, 512, 6, 8, 64, 64, 512, 2048).to(dev).eval()
for p in dec.parameters():
    if p.dim() > 1: torch.nn.init.xavier_uniform_(p)
enc = torch.randn(n, t, 512, device=dev)
ref = None
for graphs, two, fused in ((False, False, None), (True, False, None), (True, True, None), (True, True, True), (True, False, True)):
    dec.use_cuda_graphs, dec.two_streams, dec.fused_ln = graphs, two, fused
    dec._plans = {}
    out = dec.recognize_beam(enc); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); out = dec.recognize_beam(enc); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    if fused is None:
        if ref is None: ref = out
        same = torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
    else:
        same = "n/a (different kernel)"
    print(f"n={n} t={t} graphs={graphs} two_streams={two} fused_ln={fused}: {1e3 * min(ts):.2f} ms  tokens identical to eager: {same}")
