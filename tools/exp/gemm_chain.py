"""Per-kernel latency of dependent small GEMM chains in a CUDA graph (with / without PDL)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops
dev = torch.device("cuda"); ops.init()
bf = torch.bfloat16

def chain_time(M, N, K, n=64, pdl=True, out_f32=False):
    ops.set_pdl(pdl)
    a = torch.randn(M, K, device=dev).to(bf); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(bf)
    bias = torch.zeros(N, device=dev)
    o16 = torch.empty(M, N, dtype=bf, device=dev); o32 = torch.empty(M, N, dtype=torch.float32, device=dev)
    lib = ops._lib.load()
    def run():
        for _ in range(n):
            ops._lib.check(lib.sblk_gemm_fwd(a.data_ptr(), w.data_ptr(), bias.data_ptr(), None,
                                             None if out_f32 else o16.data_ptr(), o32.data_ptr() if out_f32 else None,
                                             M, N, K, 0, torch.cuda.current_stream().cuda_stream), "gemm")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        run()
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        run()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3 / n

for (M, N, K) in [(928, 512, 64), (928, 512, 512), (928, 512, 2048), (928, 1536, 512), (928, 2048, 512), (128, 64, 64)]:
    for pdl in (False, True):
        print(f"M{M} N{N} K{K} pdl={int(pdl)}: bf16-out {chain_time(M, N, K, pdl=pdl):.2f} us/kernel   "
              f"f32-out {chain_time(M, N, K, pdl=pdl, out_f32=True):.2f} us/kernel", flush=True)
