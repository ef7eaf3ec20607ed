#!/usr/bin/env python
"""L2 persistence for the encoder stack's weights: set aside part of L2 (cudaLimitPersistingL2CacheSize) and launch the
stack with an access-policy window over its (re-packed, contiguous) weights.  In the pipelined plan the encoder co-runs
with the HBM-bound layer-1 convs; the step's first phase is bound by the encoder under that contention (tools/exp/
prep_ahead_probe.py).  Plain and pipelined plans, with / without the window."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cuda.bindings import runtime as cudart
from sbl_for_multilingual_lip_reading_b200 import ops, synth
from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
from sbl_for_multilingual_lip_reading_b200.runner import PipelinedVisualEncoderPlan, VisualEncoderPlan
from sbl_for_multilingual_lip_reading_b200.video_frontend import visual_frontend
dev = torch.device("cuda"); ops.init()
N, T = 32, 29
fe = visual_frontend(None); fe.load_state_dict(synth.frontend_state_dict(1))
enc = Encoder(512, 6, 8, 64, 64, 512, 2048); enc.load_state_dict(synth.encoder_state_dict(2, 6))
fe, enc = fe.to(dev).eval(), enc.to(dev).eval()
xs = [synth.synthetic_clips(N, T, seed=7 + i).to(dev) for i in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
err, prop = cudart.cudaGetDeviceProperties(0)
print("persistingL2CacheMaxSize", prop.persistingL2CacheMaxSize >> 20, "MB; accessPolicyMaxWindowSize", prop.accessPolicyMaxWindowSize >> 20, "MB; l2", prop.l2CacheSize >> 20, "MB", flush=True)

def time_plan(plan, reps=30):
    ts = []
    for i in range(reps + 4):
        s = i % 2
        with torch.cuda.stream(plan.compute):
            plan.x[s].copy_(xs[i % 4])
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(plan.compute)
            plan.forward_device(s)
            e1.record(plan.compute)
        torch.cuda.synchronize()
        if i >= 4:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

def both(tag):
    plain = VisualEncoderPlan(fe, enc, N, T, device=dev)
    med, best = time_plan(plain); del plain
    pl = PipelinedVisualEncoderPlan(fe, enc, N, T, device=dev)
    med2, best2 = time_plan(pl); pl.close(); del pl
    print(f"{tag}: plain plan median {med:.1f} best {best:.1f} | pipelined median {med2:.1f} best {best2:.1f} ({N / med2 * 1e6:.0f} clips/s)", flush=True)

both("baseline")
# contiguous slab for the stacked weights
with torch.no_grad():
    stk = enc._get_packed().stacked
    keys = ("w_in", "w_heads", "w_fc", "w_1", "w_2")
    total = sum((stk[k].numel() * 2 + 255) // 256 * 256 for k in keys)
    slab = torch.empty(total, dtype=torch.uint8, device=dev)
    off = 0
    for k in keys:
        nb = stk[k].numel() * 2
        v = slab[off:off + nb].view(stk[k].dtype).view(stk[k].shape)
        v.copy_(stk[k]); stk[k] = v
        off += (nb + 255) // 256 * 256
print("weight slab", total >> 20, "MB", flush=True)
both("slab, no window")
for setaside_mb, ratio in ((48, 1.0), (64, 1.0)):
    e, = cudart.cudaDeviceSetLimit(cudart.cudaLimit.cudaLimitPersistingL2CacheSize, setaside_mb << 20)
    print("set-aside", setaside_mb, "MB ->", e, flush=True)
    orig = ops.encoder_stack
    def patched(*a, **kw):
        attr = cudart.cudaStreamAttrValue()
        attr.accessPolicyWindow.base_ptr = slab.data_ptr()
        attr.accessPolicyWindow.num_bytes = total
        attr.accessPolicyWindow.hitRatio = ratio
        attr.accessPolicyWindow.hitProp = cudart.cudaAccessProperty.cudaAccessPropertyPersisting
        attr.accessPolicyWindow.missProp = cudart.cudaAccessProperty.cudaAccessPropertyStreaming
        e, = cudart.cudaStreamSetAttribute(torch.cuda.current_stream().cuda_stream,
                                           cudart.cudaStreamAttrID.cudaLaunchAttributeAccessPolicyWindow, attr)
        if int(e) != 0 and not getattr(patched, "warned", False):
            print("cudaStreamSetAttribute ->", e, flush=True); patched.warned = True
        return orig(*a, **kw)
    ops.encoder_stack = patched
    import sbl_for_multilingual_lip_reading_b200.encoder as E
    both(f"window over the slab, set-aside {setaside_mb} MB")
    ops.encoder_stack = orig
