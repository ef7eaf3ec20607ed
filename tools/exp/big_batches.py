import sys, torch
sys.path.insert(0, "/root/repo")
from sbl_for_multilingual_lip_reading_b200 import ops, synth
from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
from sbl_for_multilingual_lip_reading_b200.video_frontend import Lipreading
dev = "cuda"; ops.init()
fe = Lipreading(); fe.load_state_dict(synth.frontend_state_dict(1)); fe.always_on_dropout = False
enc = Encoder(512, 6, 8, 64, 64, 512, 2048); enc.load_state_dict(synth.encoder_state_dict(2, 6))
fe, enc = fe.to(dev).eval(), enc.to(dev).eval()
with torch.no_grad():
    for n, t in ((512, 30), (256, 31), (100, 40), (33, 29), (7, 29), (1, 128)):
        x = synth.synthetic_clips(n, t, seed=3).to(dev)
        out, = enc(fe(x), [t] * n)
        k = min(n, 4)
        ref, = enc(fe(x[:k].contiguous()), [t] * k)
        torch.cuda.synchronize()
        err = ((out[:k] - ref).norm() / ref.norm()).item()
        print(f"N={n} T={t}: finite={bool(torch.isfinite(out).all())} first-{k}-clips rel diff vs small batch {err:.2e} fused={enc._use_fused_stack(n, t, False)}", flush=True)
