#!/usr/bin/env python
"""Stand-alone latency of every encoder kernel at the BASELINE shape (M = 928 tokens), back-to-back launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops
DEV = "cuda"
ops.init()
g = torch.Generator().manual_seed(0)
bf = torch.bfloat16
N, T, H = int(os.environ.get("N", 32)), 29, 8
M = N * T

def timeit(fn, n=20):
    """n back-to-back launches captured in a CUDA graph (no host launch cost) -> us per launch."""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
    st.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        for _ in range(n):
            fn()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best

x16 = torch.randn(M, 512, generator=g).to(bf).to(DEV)
h16 = torch.randn(M, 2048, generator=g).to(bf).to(DEV)
x32 = torch.randn(M, 512, generator=g).to(DEV)
w512 = (torch.randn(512, 512, generator=g) / 23).to(bf).to(DEV)
w1 = (torch.randn(2048, 512, generator=g) / 23).to(bf).to(DEV)
w2 = (torch.randn(512, 2048, generator=g) / 45).to(bf).to(DEV)
wqkv = (torch.randn(1536, 512, generator=g) / 23).to(bf).to(DEV)
b512 = torch.randn(512, generator=g).to(DEV); b2048 = torch.randn(2048, generator=g).to(DEV)
b1536 = torch.randn(1536, generator=g).to(DEV)
gm = torch.ones(512, device=DEV); bt = torch.zeros(512, device=DEV)
wh, bh = ops.pack_qkv_heads(wqkv[:512], wqkv[512:1024], wqkv[1024:], b1536[:512].contiguous(),
                            b1536[512:1024].contiguous(), b1536[1024:].contiguous(), H)
qkv16 = torch.randn(M, 1536, generator=g).to(bf).to(DEV)
for pdl in (0, 1):
    ops.set_pdl(bool(pdl))
    print(f"--- pdl={pdl} M={M}")
    print(f"cast f32->bf16 [M,512]      : {timeit(lambda: ops.cast_bf16(x32)):.2f} us")
    print(f"gemm_ln K=512 (+res)        : {timeit(lambda: ops.gemm_ln(x16, w512, gm, bt, bias=b512, residual=x32, T=T)):.2f} us")
    print(f"gemm_ln K=2048 (+res)       : {timeit(lambda: ops.gemm_ln(h16, w2, gm, bt, bias=b512, residual=x32, T=T)):.2f} us")
    print(f"linear_ln K=512 (+res)      : {timeit(lambda: ops.linear_ln(x16, w512, gm, bt, bias=b512, residual=x32, T=T)):.2f} us")
    print(f"linear_ln K=2048 (+res)     : {timeit(lambda: ops.linear_ln(h16, w2, gm, bt, bias=b512, residual=x32, T=T)):.2f} us")
    print(f"qkv_attention               : {timeit(lambda: ops.qkv_attention(x16, wh, bh, N, T, H)):.2f} us")
    print(f"gemm w1 N=2048 K=512 relu   : {timeit(lambda: ops.gemm(x16, w1, bias=b2048, relu=True, out_bf16=True)):.2f} us")
    print(f"gemm qkv N=1536 K=512       : {timeit(lambda: ops.gemm(x16, wqkv, bias=b1536, out_bf16=True)):.2f} us")
    print(f"gemm fc N=512 K=512 f32     : {timeit(lambda: ops.gemm(x16, w512, bias=b512, out_f32=True)):.2f} us")
    print(f"gemm w2 N=512 K=2048 f32    : {timeit(lambda: ops.gemm(h16, w2, bias=b512, out_f32=True)):.2f} us")
    print(f"attention                   : {timeit(lambda: ops.attention(qkv16, N, T, H)):.2f} us")
    print(f"add_layernorm               : {timeit(lambda: ops.add_layernorm(x32, gm, bt, residual=x32, T=T)):.2f} us")
