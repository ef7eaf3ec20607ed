#!/usr/bin/env python
"""Chained flat convs (sblk_flatconv3x3_chain_fwd: a ResNet stage's stride-1 convs in one launch, tile-level
dependencies through completion counters) against the same convs launched one by one: bits, then graph-replay times
(PDL on, L2 flushed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops

dev = "cuda"
ops.init()
g = torch.Generator().manual_seed(0)
bf = torch.bfloat16


def make(C, H, F, n):
    rows = ops.flat_rows(F, H, H)
    x = torch.zeros(rows, C, dtype=bf)
    grid = x.view(-1, H + 2, C)[1:].view(F, H + 1, H + 2, C)
    grid[:, :H, 1:H + 1] = torch.randn(F, H, H, C, generator=g).to(bf)
    x = ops.FlatActs(x.to(dev), F, H, H)
    ws = [ops.pack_flat_weight((torch.randn(C, 3, 3, C, generator=g) / (3 * C ** 0.5)).to(bf).to(dev)) for _ in range(n)]
    bs = [(0.1 * torch.randn(C, generator=g)).to(dev) for _ in range(n)]
    return x, ws, bs


def stage64(x, ws, bs, flags=None):
    """layer1: two BasicBlocks"""
    if flags is None:
        y1 = ops.conv3x3_flat(x, ws[0], bs[0], relu=True)
        a1 = ops.conv3x3_flat(y1, ws[1], bs[1], relu=True, residual=x)
        y2 = ops.conv3x3_flat(a1, ws[2], bs[2], relu=True)
        a2 = ops.conv3x3_flat(y2, ws[3], bs[3], relu=True, residual=a1)
        return [y1, a1, y2, a2]
    return ops.conv3x3_flat_chain(x, [(ws[0], bs[0], True, None), (ws[1], bs[1], True, "x"), (ws[2], bs[2], True, None),
                                      (ws[3], bs[3], True, 1)], flags)


def stage128(x, res, ws, bs, flags=None):
    """layer2 after its strided head: conv2 of block 0 (+ downsample branch), block 1"""
    if flags is None:
        a1 = ops.conv3x3_flat(x, ws[0], bs[0], relu=True, residual=res)
        y2 = ops.conv3x3_flat(a1, ws[1], bs[1], relu=True)
        a2 = ops.conv3x3_flat(y2, ws[2], bs[2], relu=True, residual=a1)
        return [a1, y2, a2]
    return ops.conv3x3_flat_chain(x, [(ws[0], bs[0], True, res), (ws[1], bs[1], True, None), (ws[2], bs[2], True, 0)], flags)


ok = True
for F in (() if os.environ.get('TIMING_ONLY') else (1, 2, 3, 7, 29, 120, 928)):
    x, ws, bs = make(64, 22, F, 4)
    ref = stage64(x, ws, bs)
    flags = ops.chain_flags(x, 4)
    for rep, lim in enumerate((0, 0, 6, 0, 50)):
        old = ops.set_sm_limit(lim)
        got = stage64(x, ws, bs, flags)
        ops.set_sm_limit(old)
        same = all(torch.equal(a.data, b.data) for a, b in zip(ref, got))
        ok &= same
        if not same:
            print(f"C=64 F={F} rep {rep} limit {lim}: MISMATCH", [(a.data != b.data).sum().item() for a, b in zip(ref, got)])
    x, ws, bs = make(128, 11, F, 3)
    res = make(128, 11, F, 0)[0]
    ref = stage128(x, res, ws, bs)
    flags = ops.chain_flags(x, 3)
    for rep, lim in enumerate((0, 0, 6, 0, 50)):
        old = ops.set_sm_limit(lim)
        got = stage128(x, res, ws, bs, flags)
        ops.set_sm_limit(old)
        same = all(torch.equal(a.data, b.data) for a, b in zip(ref, got))
        ok &= same
        if not same:
            print(f"C=128 F={F} rep {rep} limit {lim}: MISMATCH", [(a.data != b.data).sum().item() for a, b in zip(ref, got)])
    print(f"F={F}: done, all identical so far: {ok}", flush=True)
print("ALL IDENTICAL" if ok else "MISMATCH")

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def graph_time(fn, reps=14):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
        s.synchronize()
        gr = torch.cuda.CUDAGraph()
        old = ops.set_pdl(True)
        with torch.cuda.graph(gr, stream=s):
            fn()
        ops.set_pdl(old)
        ts = []
        for i in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); gr.replay(); e1.record(s)
            s.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]


F = 928
for lim in ((0,) if os.environ.get('TIMING_ONLY') else (0, 116)):
    ops.set_sm_limit(lim)
    x, ws, bs = make(64, 22, F, 4)
    flags = ops.chain_flags(x, 4)
    t_seq = graph_time(lambda: stage64(x, ws, bs))
    t_chain = graph_time(lambda: stage64(x, ws, bs, flags))
    t_one = graph_time(lambda: ops.conv3x3_flat(x, ws[0], bs[0], relu=True))
    print(f"sm_limit {lim}: layer1 (4 convs 64->64 H=22): one by one {t_seq:.1f} us, chain {t_chain:.1f} us, a single conv {t_one:.1f} us")
    x, ws, bs = make(128, 11, F, 3)
    res = make(128, 11, F, 0)[0]
    flags = ops.chain_flags(x, 3)
    t_seq = graph_time(lambda: stage128(x, res, ws, bs))
    t_chain = graph_time(lambda: stage128(x, res, ws, bs, flags))
    t_one = graph_time(lambda: ops.conv3x3_flat(x, ws[0], bs[0], relu=True))
    print(f"sm_limit {lim}: layer2 (3 convs 128->128 H=11): one by one {t_seq:.1f} us, chain {t_chain:.1f} us, a single conv {t_one:.1f} us")
ops.set_sm_limit(0)
