"""Does sblk_l2_prefetch help the encoder stack?  Time the one-launch stack after an L2 flush: cold, after prefetching
its weights, and fully warm (no flush)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sbl_for_multilingual_lip_reading_b200 import ops, synth
from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
ops.init()
dev = "cuda"
enc = Encoder(512, 6, 8, 64, 64, 512, 2048); enc.load_state_dict(synth.encoder_state_dict(2, 6)); enc = enc.to(dev).eval()
x16 = torch.randn(928, 512, device=dev).to(torch.bfloat16)
stk = enc._get_packed().stacked
ws = [stk[k] for k in ("w_in", "w_heads", "w_fc", "w_1", "w_2")]
out = torch.empty(928, 512, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def trial(mode):
    ts = []
    for _ in range(12):
        if mode != "warm":
            flush.zero_()
        if mode == "prefetch":
            ops.l2_prefetch(ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.encoder_stack(x16, stk, 32, 29, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
with torch.no_grad():
    for m in ("cold", "prefetch", "warm"):
        print(f"encoder stack, weights {m}: {trial(m):.1f} us")
