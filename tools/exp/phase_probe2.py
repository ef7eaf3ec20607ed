#!/usr/bin/env python
"""Timeline of the pipelined step at BASELINE configs[1] (32 x 29), graph replays with L2 flushed:
  (a) encoder stack alone (4 clusters of 8, two groups each / 8 clusters of 8),
  (b) head alone = fused stem + the first HB residual blocks on LIM SMs,   (c) both at once (gate-ordered),
  (d) tail alone = remaining blocks + pool at full width,   (e) head at full width (for reference).
python tools/exp/phase_probe2.py"""
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
feat16 = ops.cast_enc16(feat.view(N * T, 512))
shape_carrier = torch.empty((N, T, 512), dtype=torch.float32, device=dev)
gate = torch.zeros(2, dtype=torch.int32, device=dev)
side = torch.cuda.Stream(priority=-1)
state = {}


def run_enc(gpc, counter=None):
    saved = (enc.stack_cluster_size, enc._x16_override, enc._resident_counter, enc.stack_groups_per_cluster)
    enc.stack_cluster_size, enc._x16_override, enc._resident_counter, enc.stack_groups_per_cluster = 8, feat16, counter, gpc
    try:
        return enc(shape_carrier, [T] * N)[0]
    finally:
        enc.stack_cluster_size, enc._x16_override, enc._resident_counter, enc.stack_groups_per_cluster = saved


def run_blocks(a, lo, hi):
    for bi, (stride, w1, b1, w2, b2, ds) in enumerate(pk.blocks):
        if bi < lo or bi >= hi:
            continue
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
    return a


def run_head(limit, hb, fused=True):
    prev = ops.set_sm_limit(limit)
    try:
        xp = ops.raw_clip(x) if fused else ops.prep_clip(x)
        a = ops.conv3d_bn_relu_pool(xp, pk.c3w, pk.c3b, flat=True)
        state["a"] = run_blocks(a, 0, hb)
    finally:
        ops.set_sm_limit(prev)


def run_both(limit, hb, gpc):
    main = torch.cuda.current_stream()
    fork, done = torch.cuda.Event(), torch.cuda.Event()
    fork.record(main)
    side.wait_event(fork)
    with torch.cuda.stream(side):
        run_enc(gpc, gate)
        done.record(side)
    ops.gate_wait(gate, 32 if gpc == 2 else 64, 300)
    run_head(limit, hb)
    main.wait_event(done)


def run_tail(hb, limit=0):
    prev = ops.set_sm_limit(limit)
    try:
        a = run_blocks(state["a"], hb, 8)
        ops.avgpool(a, want_f32=False, want_bf16=True)
    finally:
        ops.set_sm_limit(prev)


def run_enc_with_tail(limit, hb, gpc):
    """the other split: full-width head, then the encoder next to the TAIL (layers 3-4 move few activation bytes)"""
    main = torch.cuda.current_stream()
    fork, done = torch.cuda.Event(), torch.cuda.Event()
    fork.record(main)
    side.wait_event(fork)
    with torch.cuda.stream(side):
        run_enc(gpc, gate)
        done.record(side)
    ops.gate_wait(gate, 32 if gpc == 2 else 64, 300)
    run_tail(hb, limit)
    main.wait_event(done)


def graph_time(fn, reps=20):
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


for gpc in (2, 1):
    print(f"(a) encoder stack alone, 8-CTA clusters, {gpc} group(s) per cluster: {graph_time(lambda: run_enc(gpc)):.1f} us")
for hb in (2, 3, 4, 5):
    lim = 116
    h = graph_time(lambda: run_head(lim, hb))
    hf = graph_time(lambda: run_head(0, hb))
    both = graph_time(lambda: run_both(lim, hb, 2))
    run_head(0, hb)
    torch.cuda.synchronize()
    tail = graph_time(lambda: run_tail(hb))
    print(f"head_blocks={hb}: (b) head alone on {lim} SMs {h:.1f} us, (e) on 148 SMs {hf:.1f} us, (c) encoder(2 gpc) || head {both:.1f} us, "
          f"(d) tail alone {tail:.1f} us, (c)+(d) = {both + tail:.1f} us")
for lim in (100, 108, 124, 132):
    hb = 4
    h = graph_time(lambda: run_head(lim, hb))
    both = graph_time(lambda: run_both(lim, hb, 2))
    print(f"head_blocks={hb} limit {lim}: head alone {h:.1f} us, encoder || head {both:.1f} us")

print("--- encoder next to the tail instead of the head")
for hb in (3, 4, 5):
    run_head(0, hb)
    torch.cuda.synchronize()
    hf = graph_time(lambda: run_head(0, hb))
    for lim, gpc in ((116, 2), (84, 1)):
        t_alone = graph_time(lambda: run_tail(hb, lim))
        both = graph_time(lambda: run_enc_with_tail(lim, hb, gpc))
        print(f"head_blocks={hb}: head on 148 SMs {hf:.1f} us; tail alone on {lim} SMs {t_alone:.1f} us; encoder({gpc} gpc) || tail {both:.1f} us; "
              f"sum {hf + both:.1f} us")
