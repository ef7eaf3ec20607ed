#!/usr/bin/env python
"""GPU bring-up diagnostics for libsblk (run on a B200 via gpurun).

    python tools/gpu_bringup.py <section>      sections: aux gemm conv probe conv3d

Each section runs in its own process (a trapped kernel kills the CUDA context) and compares the
hand-written kernels against torch fp32 ops evaluated on the SAME bf16-rounded operands.
torch is the checker here, never the product path.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

from sbl_for_multilingual_lip_reading_b200 import ops, _lib

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
FAILS = []


def report(name, got, ref, tol=2e-2):
    got = got.float()
    ref = ref.float()
    diff = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    rel_max = diff.max().item() / denom
    fro = (diff.norm() / (ref.norm() + 1e-12)).item()
    bad = (diff > tol * denom).float().mean().item()
    ok = fro < tol and bad < 1e-3 and bool(torch.isfinite(got).all())
    print(f"[{'OK ' if ok else 'BAD'}] {name}: rel_fro={fro:.3e} max_rel={rel_max:.3e} bad_frac={bad:.3e} "
          f"shape={tuple(got.shape)}", flush=True)
    if not ok:
        FAILS.append(name)
    return ok


def pattern(name, got, ref, rows_mod=128):
    """Where are the errors? Per row-in-tile and per column block."""
    got = got.float().reshape(-1, got.shape[-1])
    ref = ref.float().reshape(-1, ref.shape[-1])
    err = (got - ref).abs() > 0.05 * (ref.abs().max() + 1e-9)
    m, n = err.shape
    rows = err.any(dim=1)
    print(f"   {name}: bad rows {int(rows.sum())}/{m}; bad cols {int(err.any(dim=0).sum())}/{n}")
    idx = torch.nonzero(rows).flatten()[:24].tolist()
    print(f"   first bad rows: {idx}")
    rim = torch.zeros(rows_mod)
    for r in torch.nonzero(rows).flatten().tolist()[:100000]:
        rim[r % rows_mod] += 1
    print(f"   bad rows by (row % {rows_mod}) [8-bins]: {rim.reshape(8, -1).sum(1).tolist()}")
    cb = err.any(dim=0).reshape(-1, min(n, 16)).any(dim=1).int().tolist() if n % 16 == 0 else []
    print(f"   bad col blocks(16): {cb}")
    if len(idx):
        r0 = idx[0]
        print(f"   row {r0} got[:8]={got[r0, :8].tolist()}\n   row {r0} ref[:8]={ref[r0, :8].tolist()}")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def bf(x):
    return x.to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------
def sec_aux():
    g = torch.Generator(device="cpu").manual_seed(0)
    # prep
    x = torch.randn(2, 1, 5, 88, 88, generator=g).to(DEV)
    xp, _, _ = ops.prep_clip(x)
    pad = torch.zeros(2, 9, 94, 104, device=DEV)
    pad[:, 2:7, 3:91, 3:91] = x[:, 0]
    # X8[n][tp][pl][yy][x][j] = pad[n][tp][2yy+pl][2x+j]
    ref = pad.unfold(3, 8, 2)[:, :, :, :44]                      # [2,9,94,44,8]
    ref = ref.reshape(2, 9, 47, 2, 44, 8).permute(0, 1, 3, 2, 4, 5).reshape(2, 9, 2, 47 * 44 * 8)
    ref = F.pad(ref, (0, 32)).contiguous()                       # 4 zero entries per plane
    report("prep_clip", xp[:ref.numel()].reshape(ref.shape), bf(ref))
    # pack conv2d
    w = torch.randn(128, 64, 3, 3, generator=g).to(DEV)
    gam = (torch.rand(128, generator=g) + 0.5).to(DEV)
    bet = torch.randn(128, generator=g).to(DEV) * 0.1
    mu = torch.randn(128, generator=g).to(DEV) * 0.1
    var = (torch.rand(128, generator=g) + 0.5).to(DEV)
    wp, b = ops.pack_conv2d(w, gam, bet, mu, var)
    sc = gam / torch.sqrt(var + 1e-5)
    report("pack_conv2d.w", wp, bf((w * sc[:, None, None, None]).permute(0, 2, 3, 1)))
    report("pack_conv2d.b", b, bet - mu * sc, tol=1e-5)
    # pack conv3d
    w3 = torch.randn(64, 1, 5, 7, 7, generator=g).to(DEV)
    wp3, b3 = ops.pack_conv3d(w3, gam[:64].contiguous(), bet[:64].contiguous(), mu[:64].contiguous(),
                              var[:64].contiguous())
    sc3 = gam[:64] / torch.sqrt(var[:64] + 1e-5)
    ref3 = torch.zeros(64, 5, 8, 8, device=DEV)                   # [co][dt][r (2q+h)][s]
    ref3[:, :, :7, :7] = w3[:, 0] * sc3[:, None, None, None]
    report("pack_conv3d.w", wp3, bf(ref3.reshape(64, 320)))
    # cast
    a = torch.randn(1000, 512, generator=g).to(DEV)
    report("cast", ops.cast_bf16(a), bf(a))
    # avgpool
    t = bf(torch.randn(37, 3, 3, 512, generator=g)).to(DEV)
    o32, o16 = ops.avgpool(t, True, True)
    report("avgpool.f32", o32, t.float().mean(dim=(1, 2)), tol=1e-5)
    report("avgpool.bf16", o16, bf(t.float().mean(dim=(1, 2))))
    # layernorm
    M, T = 29 * 6, 29
    xx = torch.randn(M, 512, generator=g).to(DEV)
    rr = torch.randn(M, 512, generator=g).to(DEV)
    gm = torch.randn(512, generator=g).to(DEV)
    bt = torch.randn(512, generator=g).to(DEV)
    pe = torch.randn(40, 512, generator=g).to(DEV)
    lens = torch.tensor([29, 20, 1, 29, 0, 15], dtype=torch.int32, device=DEV)
    o32, o16 = ops.add_layernorm(xx, gm, bt, residual=rr, pe=pe, lengths=lens, T=T)
    ref = F.layer_norm(xx + rr, (512,), gm, bt, 1e-5) + pe[:T].repeat(6, 1)
    mask = (torch.arange(T, device=DEV)[None, :] < lens[:, None]).reshape(-1, 1).float()
    report("layernorm(res,pe,mask).f32", o32, ref * mask, tol=1e-4)
    report("layernorm.bf16", o16, bf(ref * mask))
    o32, _ = ops.add_layernorm(xx, gm, bt, T=T, want_bf16=False)
    report("layernorm(plain)", o32, F.layer_norm(xx, (512,), gm, bt, 1e-5), tol=1e-4)
    # attention
    for (N, T, lens) in [(4, 29, None), (3, 40, [40, 17, 1]), (2, 100, [100, 64]), (2, 31, None)]:
        H = 8
        qkv = bf(torch.randn(N * T, 3 * H * 64, generator=g)).to(DEV)
        lt = None if lens is None else torch.tensor(lens, dtype=torch.int32, device=DEV)
        out, probs = ops.attention(qkv, N, T, H, lengths=lt, want_probs=True)
        q, k, v = [z.float().reshape(N, T, H, 64).permute(2, 0, 1, 3).reshape(H * N, T, 64)
                   for z in qkv.split(H * 64, dim=1)]
        att = torch.bmm(q, k.transpose(1, 2)) / 8.0
        if lens is not None:
            km = (torch.arange(T, device=DEV)[None, :] >= lt[:, None])  # [N,T] True = masked key
            att = att.masked_fill(km[:, None, :].expand(N, T, T).repeat(H, 1, 1), float("-inf"))
        pr = torch.softmax(att, dim=2)
        o = torch.bmm(pr, v).reshape(H, N, T, 64).permute(1, 2, 0, 3).reshape(N * T, H * 64)
        report(f"attention N{N} T{T} lens={lens}", out, bf(o))
        report(f"attention.probs N{N} T{T}", probs, pr, tol=1e-3)


def sec_gemm():
    g = torch.Generator(device="cpu").manual_seed(1)
    cases = [(128, 64, 64), (128, 128, 128), (256, 256, 64), (928, 512, 512), (928, 1536, 512), (928, 2048, 512),
             (928, 512, 2048), (300, 192, 192), (29, 512, 512), (40000, 64, 576), (20000, 512, 1024)]
    for i, (M, N, K) in enumerate(cases):
        a = bf(torch.randn(M, K, generator=g)).to(DEV)
        w = bf(torch.randn(N, K, generator=g) / (K ** 0.5)).to(DEV)
        bias = torch.randn(N, generator=g).to(DEV)
        res = bf(torch.randn(M, N, generator=g)).to(DEV)
        ref0 = a.float() @ w.float().t()
        o16, o32 = ops.gemm(a, w, out_bf16=True, out_f32=True)
        torch.cuda.synchronize()
        ok = report(f"gemm plain M{M} N{N} K{K} f32", o32, ref0, tol=1e-4)
        report(f"gemm plain M{M} N{N} K{K} bf16", o16, bf(ref0))
        if not ok:
            pattern("gemm", o32, ref0)
        o16, o32 = ops.gemm(a, w, bias=bias, residual=res, relu=True, out_bf16=True, out_f32=True)
        ref1 = torch.relu(ref0 + bias + res.float())
        ok = report(f"gemm bias+res+relu M{M} N{N} K{K} f32", o32, ref1, tol=1e-4)
        if not ok:
            pattern("gemm.epi", o32, ref1)
    # error path: unsupported K
    try:
        ops.gemm(bf(torch.zeros(8, 48)).to(DEV), bf(torch.zeros(64, 48)).to(DEV), out_f32=True)
        print("[BAD] gemm K=48 did not raise")
        FAILS.append("gemm-raise")
    except RuntimeError as e:
        print("[OK ] gemm K=48 raises:", str(e)[:100])


def conv_ref(x, wp, bias, stride, relu, residual):
    # x NHWC bf16, wp [Co,R,S,Ci] bf16
    xr = x.float().permute(0, 3, 1, 2)
    wr = wp.float().permute(0, 3, 1, 2)
    pad = 1 if wp.shape[1] == 3 else 0
    y = F.conv2d(xr, wr, bias, stride=stride, padding=pad)
    if residual is not None:
        y = y + residual.float().permute(0, 3, 1, 2)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


def sec_conv():
    g = torch.Generator(device="cpu").manual_seed(2)
    cases = [  # F, H, Cin, Cout, R, stride
        (3, 22, 64, 64, 3, 1), (29, 22, 64, 64, 3, 1), (5, 22, 64, 128, 3, 2), (5, 22, 64, 128, 1, 2),
        (7, 11, 128, 128, 3, 1), (7, 11, 128, 256, 3, 2), (7, 11, 128, 256, 1, 2), (9, 6, 256, 256, 3, 1),
        (9, 6, 256, 512, 3, 2), (9, 6, 256, 512, 1, 2), (40, 3, 512, 512, 3, 1), (300, 22, 64, 64, 3, 1),
        (1, 22, 64, 64, 3, 1), (2, 3, 512, 512, 3, 1),
    ]
    for (Fr, H, Ci, Co, R, st) in cases:
        x = bf(torch.randn(Fr, H, H, Ci, generator=g)).to(DEV)
        wp = bf(torch.randn(Co, R, R, Ci, generator=g) / ((R * R * Ci) ** 0.5)).to(DEV)
        bias = torch.randn(Co, generator=g).to(DEV)
        out = ops.conv2d(x, wp, bias, stride=st, relu=False)
        torch.cuda.synchronize()
        ref = conv_ref(x, wp, bias, st, False, None)
        name = f"conv F{Fr} H{H} {Ci}->{Co} k{R} s{st}"
        ok = report(name, out, bf(ref))
        if not ok:
            pattern(name, out, ref)
        res = bf(torch.randn(*ref.shape, generator=g)).to(DEV)
        out = ops.conv2d(x, wp, bias, stride=st, relu=True, residual=res)
        ok = report(name + " +res+relu", out, bf(conv_ref(x, wp, bias, st, True, res)))


def sec_probe():
    """Identity-tap probe: which input pixel does each (r,s) tap of the im2col TMA actually fetch?"""
    g = torch.Generator(device="cpu").manual_seed(3)
    for (Fr, H, st, R) in [(2, 22, 1, 3), (2, 22, 2, 3), (2, 11, 2, 1), (3, 6, 1, 3)]:
        Ci = Co = 64
        x = bf(torch.randn(Fr, H, H, Ci, generator=g)).to(DEV)
        pad = 1 if R == 3 else 0
        P = (H + 2 * pad - R) // st + 1
        xpad = F.pad(x.float(), (0, 0, 4, 4, 4, 4))  # generous zero border for shift search
        for r in range(R):
            for s in range(R):
                wp = torch.zeros(Co, R, R, Ci)
                wp[torch.arange(Co), r, s, torch.arange(Ci)] = 1.0
                out = ops.conv2d(x, bf(wp).to(DEV), torch.zeros(Co, device=DEV), stride=st, relu=False).float()
                torch.cuda.synchronize()
                best = None
                for dy in range(-4, 5):
                    for dx in range(-4, 5):
                        ys = torch.arange(P) * st + dy + 4
                        xs = torch.arange(P) * st + dx + 4
                        if ys.min() < 0 or xs.min() < 0 or ys.max() >= H + 8 or xs.max() >= H + 8:
                            continue
                        cand = xpad[:, ys][:, :, xs]
                        e = (cand - out).abs().max().item()
                        if best is None or e < best[0]:
                            best = (e, dy, dx)
                exp = (r - pad, s - pad)
                ok = best[0] < 1e-6 and (best[1], best[2]) == exp
                print(f"[{'OK ' if ok else 'BAD'}] probe H{H} s{st} k{R} tap(r={r},s={s}): best shift (dy,dx)="
                      f"({best[1]},{best[2]}) err={best[0]:.3e} expected {exp}", flush=True)
                if not ok:
                    FAILS.append(f"probe H{H} s{st} tap{r}{s}")


def conv3d_ref(x, w3, gam, bet, mu, var):
    # bf16-rounded operands, fp32 math: conv3d + folded BN + relu + maxpool, output NHWC per frame
    sc = gam / torch.sqrt(var + 1e-5)
    wf = bf(w3 * sc[:, None, None, None, None]).float()
    xb = bf(x).float()
    y = F.conv3d(xb, wf, None, stride=(1, 2, 2), padding=(2, 3, 3)) + (bet - mu * sc)[None, :, None, None, None]
    y = torch.relu(y)
    y = bf(y).float()  # the kernel rounds conv outputs to bf16 before pooling
    y = F.max_pool3d(y, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    n, c, t, h, w = y.shape
    return y.permute(0, 2, 3, 4, 1).reshape(n * t, h, w, c)


def to_flat(x):
    """dense bf16 NHWC [F,H,W,C] -> ops.FlatActs (torch indexing; test helper)."""
    f, h, w, c = x.shape
    data = torch.zeros(ops.flat_rows(f, h, w), c, dtype=torch.bfloat16, device=x.device)
    v = data[(w + 2):(w + 2) + f * (h + 1) * (w + 2)].view(f, h + 1, w + 2, c)
    v[:, :h, 1:w + 1, :] = x
    return ops.FlatActs(data, f, h, w)


def halo_is_zero(fa):
    f, h, w, c = fa.f, fa.h, fa.w, fa.c
    d = fa.data.float().clone()
    v = d[(w + 2):(w + 2) + f * (h + 1) * (w + 2)].view(f, h + 1, w + 2, c)
    v[:, :h, 1:w + 1, :] = 0
    return bool((d == 0).all())


def sec_flat():
    g = torch.Generator(device="cpu").manual_seed(5)
    for (Fr, H, C) in [(1, 22, 64), (3, 22, 64), (29, 22, 64), (64, 22, 64), (7, 11, 64), (928, 22, 64),
                       (1, 11, 128), (3, 11, 128), (29, 11, 128), (64, 11, 128), (5, 6, 128), (928, 11, 128)]:
        x = bf(torch.randn(Fr, H, H, C, generator=g)).to(DEV)
        w = (torch.randn(C, C, 3, 3, generator=g) / (9 * C) ** 0.5).to(DEV)
        gam = (torch.rand(C, generator=g) + 0.5).to(DEV)
        bet = (torch.randn(C, generator=g) * 0.1).to(DEV)
        mu = (torch.randn(C, generator=g) * 0.1).to(DEV)
        var = (torch.rand(C, generator=g) + 0.5).to(DEV)
        wp, bias = ops.pack_conv2d(w, gam, bet, mu, var)
        xf = to_flat(x)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wp.float().permute(0, 3, 1, 2), bias, padding=1)
        wpf = ops.pack_flat_weight(wp)
        out = ops.conv3x3_flat(xf, wpf, bias, relu=True)
        torch.cuda.synchronize()
        ok = report(f"flatconv C{C} F{Fr} H{H} relu", out.dense(), bf(torch.relu(ref).permute(0, 2, 3, 1)))
        if not ok:
            pattern("flatconv", out.dense(), torch.relu(ref).permute(0, 2, 3, 1))
        if not halo_is_zero(out):
            print("[BAD] halo not zero"); FAILS.append("halo")
        res = bf(torch.randn(Fr, H, H, C, generator=g)).to(DEV)
        out2 = ops.conv3x3_flat(xf, wpf, bias, relu=True, residual=to_flat(res))
        report(f"flatconv C{C} F{Fr} H{H} +res+relu", out2.dense(),
               bf(torch.relu(ref + res.float().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)))
        if not halo_is_zero(out2):
            print("[BAD] halo not zero (res)"); FAILS.append("halo")
        if C == 64:
            # strided im2col conv on the flat layout == on the dense layout
            w2 = (torch.randn(128, 64, 3, 3, generator=g) / (9 * 64) ** 0.5).to(DEV)
            wp2, b2 = ops.pack_conv2d(w2)
            report(f"conv s2 on flat F{Fr} H{H}", ops.conv2d(xf, wp2, b2, stride=2), ops.conv2d(x, wp2, b2, stride=2))
    # fused conv1 + downsample (DUAL) == two separate launches, on dense and flat inputs
    for (Fr, H, Ci, Co) in [(5, 22, 64, 128), (29, 22, 64, 128), (7, 11, 128, 256), (9, 6, 256, 512), (928, 6, 256, 512)]:
        x = bf(torch.randn(Fr, H, H, Ci, generator=g)).to(DEV)
        w1 = (torch.randn(Co, Ci, 3, 3, generator=g) / (9 * Ci) ** 0.5).to(DEV)
        wd = (torch.randn(Co, Ci, 1, 1, generator=g) / Ci ** 0.5).to(DEV)
        gam = (torch.rand(Co, generator=g) + 0.5).to(DEV); bet = (torch.randn(Co, generator=g) * 0.1).to(DEV)
        mu = (torch.randn(Co, generator=g) * 0.1).to(DEV); var = (torch.rand(Co, generator=g) + 0.5).to(DEV)
        wp1, b1 = ops.pack_conv2d(w1, gam, bet, mu, var)
        wpd, bd = ops.pack_conv2d(wd, bet.abs() + 0.5, gam - 1.0, mu, var)
        y_ref = ops.conv2d(x, wp1, b1, stride=2, relu=True)
        d_ref = ops.conv2d(x, wpd, bd, stride=2, relu=False)
        y, d = ops.conv2d_dual(x, wp1, b1, wpd, bd, stride=2)
        report(f"dual conv1 F{Fr} H{H} {Ci}->{Co}", y, y_ref, tol=1e-6)
        report(f"dual ds    F{Fr} H{H} {Ci}->{Co}", d, d_ref, tol=1e-6)
        if Ci == 64:
            y2, d2 = ops.conv2d_dual(to_flat(x), wp1, b1, wpd, bd, stride=2)
            report(f"dual conv1 (flat in) F{Fr}", y2, y_ref, tol=1e-6)
            report(f"dual ds    (flat in) F{Fr}", d2, d_ref, tol=1e-6)
            P = y_ref.shape[1]
            ws = tuple(torch.zeros(ops.flat_rows(Fr, P, P), Co, dtype=torch.bfloat16, device=DEV) for _ in range(2))
            y3, d3 = ops.conv2d_dual(to_flat(x), wp1, b1, wpd, bd, stride=2, flat_ws=ws)
            report(f"dual conv1 (flat in, flat out) F{Fr}", y3.dense(), y_ref, tol=1e-6)
            report(f"dual ds    (flat in, flat out) F{Fr}", d3.dense(), d_ref, tol=1e-6)
            if not (halo_is_zero(y3) and halo_is_zero(d3)):
                print("[BAD] dual flat halo not zero"); FAILS.append("dualhalo")
    # stem flat output == dense output
    w3 = (torch.randn(64, 1, 5, 7, 7, generator=g) / (245 ** 0.5)).to(DEV)
    one, zero = torch.ones(64, device=DEV), torch.zeros(64, device=DEV)
    wp3, b3 = ops.pack_conv3d(w3, one, zero, zero, one)
    xin = torch.randn(2, 1, 7, 88, 88, generator=g).to(DEV)
    xp = ops.prep_clip(xin)
    dense = ops.conv3d_bn_relu_pool(xp, wp3, b3)
    flat = ops.conv3d_bn_relu_pool(xp, wp3, b3, flat=True)
    report("stem flat vs dense", flat.dense(), dense, tol=1e-6)
    if not halo_is_zero(flat):
        print("[BAD] stem halo not zero"); FAILS.append("stemhalo")


def sec_conv3d():
    g = torch.Generator(device="cpu").manual_seed(4)
    w3 = (torch.randn(64, 1, 5, 7, 7, generator=g) / (245 ** 0.5)).to(DEV)
    gam = (torch.rand(64, generator=g) + 0.5).to(DEV)
    bet = (torch.randn(64, generator=g) * 0.1).to(DEV)
    mu = (torch.randn(64, generator=g) * 0.1).to(DEV)
    var = (torch.rand(64, generator=g) + 0.5).to(DEV)
    wp, bias = ops.pack_conv3d(w3, gam, bet, mu, var)
    for (N, T) in [(1, 3), (1, 29), (3, 29), (2, 40), (8, 29)]:
        x = torch.randn(N, 1, T, 88, 88, generator=g).to(DEV)
        xp = ops.prep_clip(x)
        out = ops.conv3d_bn_relu_pool(xp, wp, bias)
        torch.cuda.synchronize()
        ref = conv3d_ref(x, w3, gam, bet, mu, var)
        name = f"conv3d N{N} T{T}"
        ok = report(name, out, bf(ref))
        if not ok:
            d = (out.float() - ref).abs().reshape(N * T, 22, 22, 64)
            thr = 0.05 * ref.abs().max()
            print("   bad by frame:", (d > thr).float().mean(dim=(1, 2, 3)).tolist()[:12])
            print("   bad by pooled row:", (d > thr).float().mean(dim=(0, 2, 3)).tolist())
            print("   bad by pooled col:", (d > thr).float().mean(dim=(0, 1, 3)).tolist())
            print("   bad by channel[:16]:", (d > thr).float().mean(dim=(0, 1, 2)).tolist()[:16])


def sec_enc():
    """Fused encoder kernels: gemm+LayerNorm (cluster of 4) and per-head QKV projection + attention."""
    g = torch.Generator(device="cpu").manual_seed(6)
    for (N, T, K, lens) in [(4, 29, 512, None), (32, 29, 2048, None), (3, 40, 512, [40, 17, 1]), (1, 1, 512, None),
                            (5, 7, 2048, [7, 3, 0, 7, 1]), (2, 100, 512, None)]:
        M = N * T
        a = bf(torch.randn(M, K, generator=g)).to(DEV)
        w = bf(torch.randn(512, K, generator=g) / K ** 0.5).to(DEV)
        bias = torch.randn(512, generator=g).to(DEV)
        res = torch.randn(M, 512, generator=g).to(DEV)
        gm = torch.randn(512, generator=g).to(DEV)
        bt = torch.randn(512, generator=g).to(DEV)
        pe = torch.randn(128, 512, generator=g).to(DEV)
        lt = None if lens is None else torch.tensor(lens, dtype=torch.int32, device=DEV)
        o32, o16 = ops.gemm_ln(a, w, gm, bt, bias=bias, residual=res, pe=pe, lengths=lt, T=T)
        torch.cuda.synchronize()
        ref = F.layer_norm(a.float() @ w.float().t() + bias + res, (512,), gm, bt, 1e-5) + pe[:T].repeat(N, 1)
        if lens is not None:
            ref = ref * (torch.arange(T, device=DEV)[None, :] < lt[:, None]).reshape(-1, 1).float()
        ok = report(f"gemm_ln N{N} T{T} K{K} lens={lens} f32", o32, ref, tol=2e-4)
        report(f"gemm_ln N{N} T{T} K{K} bf16", o16, bf(ref))
        if not ok:
            pattern("gemm_ln", o32, ref)
        l32, l16 = ops.linear_ln(a, w, gm, bt, bias=bias, residual=res, pe=pe, lengths=lt, T=T)
        report(f"linear_ln N{N} T{T} K{K} f32", l32, ref, tol=2e-4)
        report(f"linear_ln N{N} T{T} K{K} bf16", l16, bf(ref))
        o32, _ = ops.gemm_ln(a, w, gm, bt, T=T, want_bf16=False)
        report(f"gemm_ln plain N{N} T{T} K{K}", o32, F.layer_norm(a.float() @ w.float().t(), (512,), gm, bt, 1e-5),
               tol=2e-4)


def sec_enc2():
    g = torch.Generator(device="cpu").manual_seed(7)
    H = 8
    for (N, T, lens) in [(4, 29, None), (32, 29, None), (3, 40, [40, 17, 1]), (2, 100, [100, 64]), (1, 1, None),
                         (9, 7, None), (5, 31, [31, 1, 5, 31, 30]), (3, 64, None), (2, 65, None), (130, 1, None)]:
        K = 512
        x = bf(torch.randn(N * T, K, generator=g)).to(DEV)
        wq, wk, wv = [bf(torch.randn(H * 64, K, generator=g) / K ** 0.5).to(DEV) for _ in range(3)]
        bq, bk, bv = [torch.randn(H * 64, generator=g).to(DEV) for _ in range(3)]
        wh, bh = ops.pack_qkv_heads(wq, wk, wv, bq, bk, bv, H)
        lt = None if lens is None else torch.tensor(lens, dtype=torch.int32, device=DEV)
        out = ops.qkv_attention(x, wh, bh, N, T, H, lengths=lt)
        torch.cuda.synchronize()
        # reference = the unfused kernels' semantics: bf16-rounded projections, fp32 attention
        q, k, v = [bf(x.float() @ w_.float().t() + b_).float().reshape(N, T, H, 64).permute(2, 0, 1, 3)
                   .reshape(H * N, T, 64) for w_, b_ in ((wq, bq), (wk, bk), (wv, bv))]
        att = torch.bmm(q, k.transpose(1, 2)) / 8.0
        if lens is not None:
            km = (torch.arange(T, device=DEV)[None, :] >= lt[:, None])
            att = att.masked_fill(km[:, None, :].expand(N, T, T).repeat(H, 1, 1), float("-inf"))
        o = torch.bmm(torch.softmax(att, dim=2), v).reshape(H, N, T, 64).permute(1, 2, 0, 3).reshape(N * T, H * 64)
        ok = report(f"qkv_attention N{N} T{T} lens={lens}", out, bf(o))
        if not ok:
            pattern("qkv_attention", out, o, rows_mod=T)


def sec_stack():
    """One-launch encoder stack (cluster kernel) vs the per-step kernels and the CPU oracle.
    usage: gpu_bringup.py stack <cluster 8|16> <multicast 0|1>"""
    import os
    from oracle import visual_encoder_oracle as O
    from sbl_for_multilingual_lip_reading_b200 import synth
    from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
    cl, mc = sys.argv[2], sys.argv[3]
    os.environ["SBLK_ENC_STACK_CL"] = cl
    os.environ["SBLK_ENC_STACK_MC"] = mc
    print(f"--- encoder stack cluster={cl} multicast={mc}", flush=True)
    for nl in (1, 6):
        enc = Encoder(512, nl, 8, 64, 64, 512, 2048).eval()
        sd = synth.encoder_state_dict(2, nl)
        enc.load_state_dict(sd)
        enc = enc.to(DEV)
        g = torch.Generator(device="cpu").manual_seed(11)
        for (N, T, lens) in [(4, 29, None), (1, 1, None), (32, 29, None), (5, 29, None), (3, 40, [40, 17, 1]),
                             (7, 40, None), (2, 100, [100, 64]), (64, 31, None)]:
            x = torch.randn(N, T, 512, generator=g)
            ln = lens if lens is not None else [T] * N
            with torch.no_grad():
                enc.fused_stack = True
                got, = enc(x.to(DEV), ln)
                torch.cuda.synchronize()
                enc.fused_stack = False
                ref, = enc(x.to(DEV), ln)
                torch.cuda.synchronize()
                report(f"stack L{nl} N{N} T{T} lens={lens} vs per-step kernels", got, ref, tol=1e-2)
                if N * T <= 400:
                    orc = O.encoder_forward(x, ln, {k: v for k, v in sd.items()}, n_layers=nl)[0]
                    report(f"stack L{nl} N{N} T{T} lens={lens} vs oracle", got.cpu(), orc, tol=2e-2)
    x = torch.randn(32, 29, 512, generator=g).to(DEV)
    ln = [29] * 32
    with torch.no_grad():
        for fused in (True, False):
            enc.fused_stack = fused
            for pdl in (False, True):
                ops.set_pdl(pdl)
                ms = timeit(lambda: enc(x, ln))
                print(f"perf encoder N32 T29 L6 fused={fused} pdl={pdl}: {ms * 1e3:.1f} us", flush=True)
    ops.set_pdl(False)


def sec_perf():
    """Quick per-layer timing at the C2 shape (F = 928) to see where the time goes."""
    g = torch.Generator(device="cpu").manual_seed(5)
    Fr = 928

    layers = [(22, 64, 64, 3, 1), (22, 64, 128, 3, 2), (11, 128, 128, 3, 1), (22, 64, 128, 1, 2),
              (11, 128, 256, 3, 2), (6, 256, 256, 3, 1), (6, 256, 512, 3, 2), (3, 512, 512, 3, 1)]
    for (H, Ci, Co, R, st) in layers:
        x = bf(torch.randn(Fr, H, H, Ci, generator=g)).to(DEV)
        wp = bf(torch.randn(Co, R, R, Ci, generator=g) / ((R * R * Ci) ** 0.5)).to(DEV)
        bias = torch.zeros(Co, device=DEV)
        pad = 1 if R == 3 else 0
        P = (H + 2 * pad - R) // st + 1
        out = torch.empty(Fr, P, P, Co, dtype=torch.bfloat16, device=DEV)
        ms = timeit(lambda: ops.conv2d(x, wp, bias, stride=st, relu=True, out=out))
        fl = 2.0 * Fr * P * P * Co * R * R * Ci
        print(f"perf conv H{H} {Ci}->{Co} k{R} s{st}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    for (M, N, K) in [(928, 1536, 512), (928, 512, 512), (928, 2048, 512), (928, 512, 2048)]:
        a = bf(torch.randn(M, K, generator=g)).to(DEV)
        w = bf(torch.randn(N, K, generator=g)).to(DEV)
        ms = timeit(lambda: ops.gemm(a, w, out_bf16=True))
        print(f"perf gemm M{M} N{N} K{K}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    for (H, Ci, Co) in [(22, 64, 128), (11, 128, 256), (6, 256, 512)]:
        x = bf(torch.randn(928, H, H, Ci, generator=g)).to(DEV)
        wa = bf(torch.randn(Co, 3, 3, Ci, generator=g) / (9 * Ci) ** 0.5).to(DEV)
        wb = bf(torch.randn(Co, 1, 1, Ci, generator=g) / Ci ** 0.5).to(DEV)
        bz = torch.zeros(Co, device=DEV)
        ms = timeit(lambda: ops.conv2d_dual(x, wa, bz, wb, bz, stride=2))
        print(f"perf dual conv3x3+ds H{H} {Ci}->{Co} s2: {ms * 1e3:.1f} us", flush=True)
    xd = bf(torch.randn(928, 22, 22, 64, generator=g)).to(DEV)
    xf = to_flat(xd)
    wf = ops.pack_flat_weight(bf(torch.randn(64, 3, 3, 64, generator=g) / 24).to(DEV))
    bfz = torch.zeros(64, device=DEV)
    of = torch.empty_like(xf.data)
    ms = timeit(lambda: ops.conv3x3_flat(xf, wf, bfz, relu=True, residual=xf, out=of))
    print(f"perf flatconv H22 64->64 (+res): {ms * 1e3:.1f} us  {2.0 * 928 * 484 * 64 * 576 / ms / 1e9:.1f} TFLOP/s",
          flush=True)
    ms = timeit(lambda: ops.conv3x3_flat(xf, wf, bfz, relu=True, out=of))
    print(f"perf flatconv H22 64->64 (no res): {ms * 1e3:.1f} us  {2.0 * 928 * 484 * 64 * 576 / ms / 1e9:.1f} TFLOP/s",
          flush=True)
    xd2 = bf(torch.randn(928, 11, 11, 128, generator=g)).to(DEV)
    xf2 = to_flat(xd2)
    wf2 = ops.pack_flat_weight(bf(torch.randn(128, 3, 3, 128, generator=g) / 34).to(DEV))
    bfz2 = torch.zeros(128, device=DEV)
    of2 = torch.empty_like(xf2.data)
    ms = timeit(lambda: ops.conv3x3_flat(xf2, wf2, bfz2, relu=True, residual=xf2, out=of2))
    print(f"perf flatconv H11 128->128 (+res): {ms * 1e3:.1f} us  {2.0 * 928 * 121 * 128 * 1152 / ms / 1e9:.1f} TFLOP/s",
          flush=True)
    ms = timeit(lambda: ops.conv3x3_flat(xf2, wf2, bfz2, relu=True, out=of2))
    print(f"perf flatconv H11 128->128 (no res): {ms * 1e3:.1f} us  {2.0 * 928 * 121 * 128 * 1152 / ms / 1e9:.1f} TFLOP/s",
          flush=True)
    w3 = (torch.randn(64, 1, 5, 7, 7, generator=g) / 16).to(DEV)
    one = torch.ones(64, device=DEV)
    zero = torch.zeros(64, device=DEV)
    wp3, b3 = ops.pack_conv3d(w3, one, zero, zero, one)
    x = torch.randn(32, 1, 29, 88, 88, generator=g).to(DEV)
    xp = ops.prep_clip(x)
    out = torch.empty(928, 22, 22, 64, dtype=torch.bfloat16, device=DEV)
    ms = timeit(lambda: ops.conv3d_bn_relu_pool(xp, wp3, b3, out=out))
    print(f"perf conv3d N32 T29: {ms * 1e3:.1f} us  {2.0 * 928 * 44 * 44 * 64 * 245 / ms / 1e9:.1f} TFLOP/s")
    ms = timeit(lambda: ops.prep_clip(x, out=xp[0]))
    print(f"perf prep N32 T29: {ms * 1e3:.1f} us")


if __name__ == "__main__":
    sec = sys.argv[1]
    t0 = time.time()
    print(f"=== section {sec}: SMs={ops.init()} torch={torch.__version__} dev={torch.cuda.get_device_name(0)}",
          flush=True)
    if len(sys.argv) > 2 and sys.argv[2] == "pdl":
        ops.set_pdl(True)
        print("PDL enabled")
    try:
        {"aux": sec_aux, "gemm": sec_gemm, "conv": sec_conv, "probe": sec_probe, "conv3d": sec_conv3d, "flat": sec_flat, "enc": sec_enc, "enc2": sec_enc2,
         "perf": sec_perf, "stack": sec_stack}[sec]()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        import traceback
        traceback.print_exc()
        print(f"EXCEPTION in section {sec}: {e}; watchdog=0x{_lib.load().sblk_watchdog_code():08x}", flush=True)
        FAILS.append(f"exception:{sec}")
    print(f"=== section {sec} done in {time.time() - t0:.1f}s; failures: {FAILS}", flush=True)
    sys.exit(1 if FAILS else 0)
