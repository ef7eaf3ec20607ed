#!/usr/bin/env python
"""Encoder stack: two clip groups interleaved per cluster (groups_per_cluster=2) against one — bit-exactness, timings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops, synth
from sbl_for_multilingual_lip_reading_b200.encoder import Encoder

dev = torch.device("cuda")
ops.init()
L = 6
enc = Encoder(512, L, 8, 64, 64, 512, 2048)
enc.load_state_dict(synth.encoder_state_dict(2, L))
enc = enc.to(dev).eval()
ok = True
with torch.no_grad():
    stk = enc._get_packed().stacked
    for (N, T, ragged) in [(32, 29, False), (32, 29, True), (28, 29, False), (5, 29, True), (1, 29, False), (9, 40, True),
                           (7, 100, False), (3, 1, False), (48, 29, False)]:
        g = torch.Generator().manual_seed(N * 1000 + T)
        x16 = torch.randn(N * T, 512, generator=g).to(dev).to(ops.enc16_dtype())
        lengths = None
        if ragged:
            lengths = torch.randint(1, T + 1, (N,), generator=g).to(torch.int32).to(dev)
        outs = {}
        for (cl, gpc) in [(8, 1), (8, 2), (16, 1), (16, 2)]:
            groups = -(-N // max(1, 128 // T))
            clusters = -(-groups // gpc)
            if cl == 16 and clusters > 7:
                continue
            outs[(cl, gpc)] = ops.encoder_stack(x16, stk, N, T, lengths=lengths, cluster_size=cl, groups_per_cluster=gpc).clone()
        torch.cuda.synchronize()
        ref = outs[(8, 1)]
        line = f"N={N} T={T} ragged={ragged}:"
        for k, o in outs.items():
            same = torch.equal(o, ref)
            ok &= same
            line += f"  cl={k[0]} gpc={k[1]} {'==' if same else 'DIFF ' + str((o - ref).abs().max().item())}"
        print(line, flush=True)

    # timings, L2 flushed
    N, T = 32, 29
    x16 = torch.randn(N * T, 512).to(dev).to(ops.enc16_dtype())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = torch.empty(N * T, 512, device=dev)
    for (cl, gpc) in [(8, 1), (8, 2), (16, 2)]:
        ts = []
        for it in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.encoder_stack(x16, stk, N, T, cluster_size=cl, groups_per_cluster=gpc, out=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts = sorted(ts[2:])
        clusters = -(-8 // gpc)
        print(f"N=32 T=29 cl={cl} gpc={gpc} ({clusters * cl} CTAs): {ts[len(ts) // 2]:.1f} us (cold L2)", flush=True)
print("ALL OK" if ok else "MISMATCH")
