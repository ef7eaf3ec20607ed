#!/usr/bin/env python
"""Small fixed workloads for ncu: python tools/prof_target.py <conv3d|l1conv|l4conv|forward> [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops, synth

dev = "cuda"
what = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ops.init()
g = torch.Generator().manual_seed(0)
bf = torch.bfloat16
if what == "conv3d":
    w3 = (torch.randn(64, 1, 5, 7, 7, generator=g) / 16).to(dev)
    one, zero = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    wp, b = ops.pack_conv3d(w3, one, zero, zero, one)
    x = synth.synthetic_clips(32, 29, seed=7).to(dev)
    xp = ops.prep_clip(x)
    out = torch.empty(928, 22, 22, 64, dtype=bf, device=dev)
    for _ in range(iters):
        ops.conv3d_bn_relu_pool(xp, wp, b, out=out)
elif what == "stemfused":   # the stem with the clip prep fused in (fp32 clips in, no prepped copy)
    w3 = (torch.randn(64, 1, 5, 7, 7, generator=g) / 16).to(dev)
    one, zero = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    wp, b = ops.pack_conv3d(w3, one, zero, zero, one)
    x = synth.synthetic_clips(32, 29, seed=7).to(dev)
    out = ops.conv3d_bn_relu_pool(ops.raw_clip(x), wp, b, flat=True)
    for _ in range(iters):
        ops.conv3d_bn_relu_pool(ops.raw_clip(x), wp, b, out=out.data, flat=True)
elif what in ("l1conv", "l2conv", "l3conv", "l4conv"):
    H, C = {"l1conv": (22, 64), "l2conv": (11, 128), "l3conv": (6, 256), "l4conv": (3, 512)}[what]
    x = torch.randn(928, H, H, C, generator=g).to(bf).to(dev)
    w = (torch.randn(C, 3, 3, C, generator=g) / (9 * C) ** 0.5).to(bf).to(dev)
    bias = torch.zeros(C, device=dev)
    out = torch.empty(928, H, H, C, dtype=bf, device=dev)
    for _ in range(iters):
        ops.conv2d(x, w, bias, relu=True, residual=x, out=out)
if what == "flat":
    F_, H = 928, 22
    rows = ops.flat_rows(F_, H, H)
    data = torch.zeros(rows, 64, dtype=bf, device=dev)
    v = data[(H + 2):(H + 2) + F_ * (H + 1) * (H + 2)].view(F_, H + 1, H + 2, 64)
    v[:, :H, 1:H + 1, :] = torch.randn(F_, H, H, 64, generator=g).to(bf).to(dev)
    xf = ops.FlatActs(data, F_, H, H)
    w = ops.pack_flat_weight((torch.randn(64, 3, 3, 64, generator=g) / 24).to(bf).to(dev))
    bias = torch.zeros(64, device=dev)
    out = torch.empty_like(data)
    for _ in range(iters):
        ops.conv3x3_flat(xf, w, bias, relu=True, residual=xf, out=out)
if what == "forward":
    # whole hot path, eager (one launch per kernel), BASELINE configs[1] shape: 32 clips x 29 frames
    from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
    from sbl_for_multilingual_lip_reading_b200.video_frontend import visual_frontend
    fe = visual_frontend(None)
    fe.load_state_dict(synth.frontend_state_dict(1))
    enc = Encoder(512, 6, 8, 64, 64, 512, 2048)
    enc.load_state_dict(synth.encoder_state_dict(2, 6))
    fe, enc = fe.to(dev).eval(), enc.to(dev).eval()
    x = synth.synthetic_clips(32, 29, seed=7).to(dev)
    with torch.no_grad():
        for _ in range(iters):
            out, = enc(fe(x), [29] * 32)
    torch.cuda.synchronize()
    print("done forward", float(out.abs().mean()))
if what == "dual4":
    # strided block head of layer4: conv3x3 s2 + 1x1 downsample in one pass (6x6x256 -> 3x3x512)
    x = torch.randn(928, 6, 6, 256, generator=g).to(bf).to(dev)
    wa = (torch.randn(512, 3, 3, 256, generator=g) / (9 * 256) ** 0.5).to(bf).to(dev)
    wb = (torch.randn(512, 1, 1, 256, generator=g) / 16).to(bf).to(dev)
    bz = torch.zeros(512, device=dev)
    for _ in range(iters):
        ops.conv2d_dual(x, wa, bz, wb, bz, stride=2)
if what == "dual2":
    # strided block head of layer2: conv3x3 s2 + 1x1 downsample in one pass (22x22x64 -> 11x11x128)
    x = torch.randn(928, 22, 22, 64, generator=g).to(bf).to(dev)
    wa = (torch.randn(128, 3, 3, 64, generator=g) / 24).to(bf).to(dev)
    wb = (torch.randn(128, 1, 1, 64, generator=g) / 8).to(bf).to(dev)
    bz = torch.zeros(128, device=dev)
    for _ in range(iters):
        ops.conv2d_dual(x, wa, bz, wb, bz, stride=2)
if what in ("fold3", "fold4"):
    # residual block head with the downsample branch folded into conv2 (layer3.0: 11x11x128 -> 6x6x256,
    # layer4.0: 6x6x256 -> 3x3x512): conv1 3x3/s2 on 256-wide pair tiles, then conv2 + K-extension
    cin, cout, h = (128, 256, 11) if what == "fold3" else (256, 512, 6)
    x = torch.randn(928, h, h, cin, generator=g).to(bf).to(dev)
    w1 = (torch.randn(cout, 3, 3, cin, generator=g) / (9 * cin) ** 0.5).to(bf).to(dev)
    w2 = (torch.randn(cout, 3, 3, cout, generator=g) / (9 * cout) ** 0.5).to(bf).to(dev)
    wd = (torch.randn(cout, 1, 1, cin, generator=g) / cin ** 0.5).to(bf).to(dev)
    bz = torch.zeros(cout, device=dev)
    for _ in range(iters):
        y = ops.conv2d(x, w1, bz, stride=2)
        ops.conv2d(y, w2, bz, stride=1, ext=(x, wd, 2))
if what in ("block3a", "block3b"):
    # a whole layer-3 residual block as one launch (igemm2_block_kernel): head block (11x11x128 -> 6x6x256, stride 2 +
    # folded downsample branch) or identity block (6x6x256)
    cin, h, stride = (128, 11, 2) if what == "block3a" else (256, 6, 1)
    x = torch.randn(928, h, h, cin, generator=g).to(bf).to(dev)
    w1 = (torch.randn(256, 3, 3, cin, generator=g) / (9 * cin) ** 0.5).to(bf).to(dev)
    w2 = (torch.randn(256, 3, 3, 256, generator=g) / (9 * 256) ** 0.5).to(bf).to(dev)
    wd = (torch.randn(256, 1, 1, cin, generator=g) / cin ** 0.5).to(bf).to(dev) if stride == 2 else None
    bz = torch.zeros(256, device=dev)
    for _ in range(iters):
        ops.conv_block(x, w1, bz, w2, bz, w_ds=wd, stride=stride)
if what == "stack":
    # the one-launch encoder stack alone, BASELINE configs[1] shape
    from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
    enc = Encoder(512, 6, 8, 64, 64, 512, 2048)
    enc.load_state_dict(synth.encoder_state_dict(2, 6))
    enc = enc.to(dev).eval()
    # SBLK_PROF_STACK="cluster,groups_per_cluster" (e.g. "8,2": the form the pipelined plan runs on 32 SMs)
    cfg = os.environ.get("SBLK_PROF_STACK")
    if cfg:
        enc.stack_cluster_size, enc.stack_groups_per_cluster = (int(v) for v in cfg.split(","))
    feat = torch.randn(32, 29, 512, generator=g).to(dev)
    with torch.no_grad():
        for _ in range(iters):
            out, = enc(feat, [29] * 32)
torch.cuda.synchronize()
print("done", what)
