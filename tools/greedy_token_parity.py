#!/usr/bin/env python
"""Greedy-decode token parity on 1,000 synthetic clips (BASELINE.json north_star), in two steps because the
reference decoder exists only in the build container and the B200 only on the GPU box:

  GPU box :  python tools/greedy_token_parity.py export     -> gpurun_out/enc_out_cuda_1000.npy (CUDA encoder outputs)
  here    :  python tools/greedy_token_parity.py compare    -> profiles/r01_greedy_token_parity.json

`compare` runs the UNMODIFIED reference frontend + encoder (fp32, CPU) on the same 1,000 structured clips
(synth.structured_clips(40, 29, seed=5000 + c), c = 0..24), then the reference SBL bidirectional decoder's greedy
search `Decoder.recognize_beam` (transformer/decoder.py:301-385, what `Transformer.recognize` / test.py:163-174 call)
once on the reference encoder outputs and once on the CUDA encoder outputs, and compares the emitted l2r / r2l token
sequences clip by clip.  Weights: synth frontend / encoder state dicts; decoder as `Transformer.__init__` leaves it
(xavier_uniform_, transformer.py:18-20) under torch.manual_seed(7).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sbl_for_multilingual_lip_reading_b200 import synth  # noqa: E402

CHUNK, CHUNKS, T = 40, 25, 29
OUT = os.path.join(ROOT, "gpurun_out", "enc_out_cuda_1000.npy")


def export():
    from sbl_for_multilingual_lip_reading_b200 import ops
    from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
    from sbl_for_multilingual_lip_reading_b200.video_frontend import Lipreading
    dev = torch.device("cuda:0")
    ops.init()
    fe = Lipreading(); fe.load_state_dict(synth.frontend_state_dict(1)); fe.always_on_dropout = False
    enc = Encoder(512, 6, 8, 64, 64, 512, 2048); enc.load_state_dict(synth.encoder_state_dict(2, 6))
    fe, enc = fe.to(dev).eval(), enc.to(dev).eval()
    outs = []
    with torch.no_grad():
        for c in range(CHUNKS):
            x = synth.structured_clips(CHUNK, T, seed=5000 + c).to(dev)
            out, = enc(fe(x), [T] * CHUNK)
            outs.append(out.cpu())
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.save(OUT, torch.cat(outs).numpy())
    print("saved", OUT, torch.cat(outs).shape)


def compare():
    ref_dir = "/root/reference/SBL_Multilingual_Lip_reading"
    sys.path.insert(0, ref_dir)
    from transformer.decoder import Decoder
    from transformer.encoder import Encoder
    from transformer.transformer import Transformer
    torch.set_grad_enabled(False)
    torch.manual_seed(7)
    enc = Encoder(512, 6, 8, 64, 64, 512, 2048, dropout=0.1, pe_maxlen=5000)
    dec = Decoder(0, 1, 58, 512, 6, 8, 64, 64, 512, 2048, dropout=0.1, tgt_emb_prj_weight_sharing=1, pe_maxlen=5000)
    model = Transformer(enc, dec, None)
    sd = dict(synth.frontend_state_dict(1, prefix="visual_frontend."))
    sd.update(synth.encoder_state_dict(2, 6, prefix="encoder."))
    model.load_state_dict(sd, strict=False)
    model.eval()
    cuda_out = torch.from_numpy(np.load(OUT))
    same_l2r = same_r2l = total = 0
    ctl_l2r = ctl_r2l = 0   # control: the reference's own encoder output rounded ONCE to bf16 (relative error ~2e-3)
    worst = 0.0
    distinct = set()
    for c in range(CHUNKS):
        x = synth.structured_clips(CHUNK, T, seed=5000 + c)
        feat = model.visual_frontend._frontend_forward(x).view(CHUNK, T, 512)   # forward() minus the always-on dropout
        ref_out, *_ = model.encoder(feat, [T] * CHUNK)
        got = cuda_out[c * CHUNK:(c + 1) * CHUNK]
        worst = max(worst, ((got - ref_out).norm() / ref_out.norm()).item())
        a_l2r, a_r2l = model.decoder.recognize_beam(ref_out)
        b_l2r, b_r2l = model.decoder.recognize_beam(got)
        c_l2r, c_r2l = model.decoder.recognize_beam(ref_out.to(torch.bfloat16).float())
        for i in range(CHUNK):
            same_l2r += int(torch.equal(a_l2r[i], b_l2r[i]))
            same_r2l += int(torch.equal(a_r2l[i], b_r2l[i]))
            ctl_l2r += int(torch.equal(a_l2r[i], c_l2r[i]))
            ctl_r2l += int(torch.equal(a_r2l[i], c_r2l[i]))
            distinct.add(tuple(a_l2r[i].tolist()))
        total += CHUNK
        print(f"chunk {c + 1}/{CHUNKS}: identical l2r {same_l2r}/{total}, r2l {same_r2l}/{total}", flush=True)
    res = {"clips": total, "identical_l2r_sequences": same_l2r, "identical_r2l_sequences": same_r2l,
           "control_bf16_rounded_reference_output": {"identical_l2r_sequences": ctl_l2r,
                                                     "identical_r2l_sequences": ctl_r2l},
           "distinct_reference_l2r_sequences": len(distinct), "max_chunk_rel_fro_error_encoder_output": worst,
           "decoder": "reference Decoder.recognize_beam (greedy, maxlen 16), transformer/decoder.py:301-385",
           "note": "with random weights the greedy decoder emits (nearly) clip-independent tokens (SURVEY.md 8c caveat 3)"}
    with open(os.path.join(ROOT, "profiles", "r01_greedy_token_parity.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    {"export": export, "compare": compare}[sys.argv[1]]()
