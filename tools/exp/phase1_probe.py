#!/usr/bin/env python
"""Where does the first phase of the pipelined step go?  Graph-replay times (L2 flushed) at BASELINE configs[1] of
  (a) the encoder stack alone (8-CTA clusters, bf16 input),   (b) clip prep + stem alone on 148 / 84 SMs,
  (c) both at once (gate-ordered), (d) the trunk alone (layer1..4 + pool) at full width.
python tools/exp/phase1_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops, synth
from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
from sbl_for_multilingual_lip_reading_b200.video_frontend import visual_frontend

dev = torch.device("cuda")
ops.init()
N, T = 32, 29
fe = visual_frontend(None); fe.load_state_dict(synth.frontend_state_dict(1))
enc = Encoder(512, 6, 8, 64, 64, 512, 2048); enc.load_state_dict(synth.encoder_state_dict(2, 6))
fe, enc = fe.to(dev).eval(), enc.to(dev).eval()
x = synth.synthetic_clips(N, T, seed=7).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pk = fe._get_packed()
with torch.no_grad():
    feat = fe(x)
feat16 = ops.cast_bf16(feat.view(N * T, 512))
shape_carrier = torch.empty((N, T, 512), dtype=torch.float32, device=dev)
gate = torch.zeros(2, dtype=torch.int32, device=dev)
side = torch.cuda.Stream(priority=-1)
stem_out = {}


def run_enc(counter=None):
    enc.stack_cluster_size, enc._x16_override, enc._resident_counter = 8, feat16, counter
    try:
        return enc(shape_carrier, [T] * N)[0]
    finally:
        enc.stack_cluster_size, enc._x16_override, enc._resident_counter = 0, None, None


def run_head(limit):
    prev = ops.set_sm_limit(limit)
    try:
        xp = ops.prep_clip(x)
        stem_out["a"] = ops.conv3d_bn_relu_pool(xp, pk.c3w, pk.c3b, flat=True)
    finally:
        ops.set_sm_limit(prev)


def run_both(limit, use_gate=True):
    main = torch.cuda.current_stream()
    fork, done = torch.cuda.Event(), torch.cuda.Event()
    fork.record(main)
    side.wait_event(fork)
    with torch.cuda.stream(side):
        run_enc(gate if use_gate else None)
        done.record(side)
    if use_gate:
        ops.gate_wait(gate, 64, 300)
    run_head(limit)
    main.wait_event(done)


def run_trunk():
    # the frontend chain from layer1 on: reuse the chain body with a pre-computed stem output
    a = stem_out["a"]
    for bi, (stride, w1, b1, w2, b2, ds) in enumerate(pk.blocks):
        if isinstance(a, ops.FlatActs) and stride == 1 and ds is None:
            y = ops.conv3x3_flat(a, w1, b1, relu=True)
            a = ops.conv3x3_flat(y, w2, b2, relu=True, residual=a)
            continue
        if ds is not None:
            if w2.dim() == 2:
                y, res = ops.conv2d_dual(a, w1, b1, ds[0], ds[1], stride=stride, relu=True,
                                         flat_ws=fe._flat_workspace(a, w1.shape[0], stride, 0))
                a = ops.conv3x3_flat(y, w2, b2, relu=True, residual=res)
                continue
            y, res = ops.conv2d_dual(a, w1, b1, ds[0], ds[1], stride=stride, relu=True)
        else:
            y, res = ops.conv2d(a, w1, b1, stride=stride, relu=True), a
        a = ops.conv2d(y, w2, b2, stride=1, relu=True, residual=res)
    ops.avgpool(a, want_f32=False, want_bf16=True)


def graph_time(fn, reps=30):
    s = torch.cuda.Stream()
    with torch.no_grad(), torch.cuda.stream(s):
        for _ in range(2):
            fn()
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g, stream=s):
        fn()
    ts = []
    for _ in range(reps):
        with torch.cuda.stream(s):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); g.replay(); e1.record(s)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


print(f"(a) encoder stack alone, 8-CTA clusters:        {graph_time(run_enc):.1f} us")
for lim in (0, 84, 74):
    print(f"(b) prep + stem alone, SM limit {lim:3d}:            {graph_time(lambda: run_head(lim)):.1f} us")
for lim in (84, 74):
    print(f"(c) encoder || prep + stem on {lim} SMs, gate:      {graph_time(lambda: run_both(lim)):.1f} us")
print(f"(c) encoder || prep + stem on 84 SMs, no gate:   {graph_time(lambda: run_both(84, False)):.1f} us")
run_head(0)
torch.cuda.synchronize()
print(f"(d) trunk alone (layer1..4 + pool), 148 SMs:     {graph_time(run_trunk):.1f} us")
