#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

It imports the unmodified reference modules from /root/reference/SBL_Multilingual_Lip_reading, loads the
seeded synthetic state dicts of `sbl_for_multilingual_lip_reading_b200.synth` into them, runs them on CPU in
fp32 / eval mode and stores inputs' seeds and outputs.  The always-on dropout of Lipreading.forward
(transformer/video_frontend.py:122) is avoided by calling `_frontend_forward` (:111-117), which is
everything forward() does apart from that dropout and a view.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/SBL_Multilingual_Lip_reading"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from transformer.encoder import Encoder as RefEncoder  # noqa: E402
from transformer.video_frontend import Lipreading as RefLipreading  # noqa: E402
from transformer.decoder import Decoder as RefDecoder  # noqa: E402
from transformer.transformer import Transformer as RefTransformer  # noqa: E402

from sbl_for_multilingual_lip_reading_b200 import synth  # noqa: E402

torch.set_grad_enabled(False)
torch.manual_seed(0)


def ref_frontend(seed=1):
    m = RefLipreading()
    m.load_state_dict(synth.frontend_state_dict(seed))
    return m.eval()


def ref_encoder(n_layers, seed=2):
    m = RefEncoder(512, n_layers, 8, 64, 64, 512, 2048, dropout=0.1, pe_maxlen=5000)
    m.load_state_dict(synth.encoder_state_dict(seed, n_layers))
    return m.eval()


def main():
    out = {}
    fe = ref_frontend()

    # -- state-dict contract of the reference classes (keys, shapes, dtypes) and of the SBL Transformer
    enc6 = ref_encoder(6)
    contract = {
        "Lipreading": {k: [list(v.shape), str(v.dtype)] for k, v in fe.state_dict().items()},
        "Encoder6": {k: [list(v.shape), str(v.dtype)] for k, v in enc6.state_dict().items()},
    }
    dec = RefDecoder(0, 1, 58, 512, 6, 8, 64, 64, 512, 2048, dropout=0.1, tgt_emb_prj_weight_sharing=1,
                     pe_maxlen=5000)
    tr = RefTransformer(RefEncoder(512, 6, 8, 64, 64, 512, 2048, dropout=0.1, pe_maxlen=5000), dec, None)
    contract["Transformer_hot_path_keys"] = sorted(
        k for k in tr.state_dict().keys() if k.startswith("visual_frontend.") or k.startswith("encoder."))
    contract["Transformer_num_keys"] = len(tr.state_dict())
    with open(os.path.join(HERE, "state_dict_contract.json"), "w") as f:
        json.dump(contract, f, indent=0, sort_keys=True)

    # -- frontend3D alone (Conv3d+BN+ReLU+MaxPool3d), N=1, T=2 -> [1,64,2,22,22]
    x = synth.synthetic_clips(1, 2, seed=11)
    out["frontend3d_T2"] = fe.frontend3D(x).numpy()

    # -- first BasicBlock and first strided block on the pooled stem output, T=2
    y = fe.frontend3D(x).transpose(1, 2).contiguous().view(-1, 64, 22, 22)
    out["layer1_0_T2"] = fe.resnet18.layer1[0](y.clone()).numpy()
    out["layer2_0_T2"] = fe.resnet18.layer2[0](fe.resnet18.layer1(y.clone())).numpy()

    # -- BASELINE config 1: 1 LRW clip 29x88x88, frontend + trunk -> [29,512]
    x = synth.synthetic_clips(1, 29, seed=7)
    out["frontend_c1"] = fe._frontend_forward(x).numpy()

    # -- zero-padded trailing frame as the SBL loader produces (T=30, last frame zero), N=2, T=6 here
    x = synth.synthetic_clips(2, 6, seed=8, pad_frames=1)
    feat = fe._frontend_forward(x)
    out["frontend_N2_T6_pad1"] = feat.numpy()

    # -- encoder, full lengths (every reference call site), 6 and 3 layers
    g = torch.Generator().manual_seed(21)
    xin = torch.randn(1, 29, 512, generator=g)
    out["encoder6_N1_T29"] = enc6(xin, [29])[0].numpy()
    enc3 = ref_encoder(3, seed=3)
    xin3 = torch.randn(2, 31, 512, generator=g)
    out["encoder3_N2_T31"] = enc3(xin3, [31, 31])[0].numpy()

    # -- encoder, ragged lengths + return_attns (signature features the reference never exercises)
    xin_r = torch.randn(3, 12, 512, generator=g)
    eo, attns = enc6(xin_r, [12, 7, 1], return_attns=True)
    out["encoder6_ragged_out"] = eo.numpy()
    out["encoder6_ragged_attn0"] = attns[0].numpy()
    out["encoder6_ragged_attn5"] = attns[5].numpy()

    # -- whole hot path as Transformer.forward drives it (lengths = [T]*N), N=2, T=6, 6 layers
    out["visual_encoder_N2_T6"] = enc6(feat.view(2, 6, 512), [6, 6])[0].numpy()

    # -- T = 40 (LRW-1000-shaped), N=1: encoder only (frontend is T-agnostic per frame + temporal conv)
    xin40 = torch.randn(1, 40, 512, generator=g)
    out["encoder6_N1_T40"] = enc6(xin40, [40])[0].numpy()

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    for k, v in out.items():
        print(f"{k}: shape={v.shape} mean={v.mean():+.6f} std={v.std():.6f}")
    # encoder inputs are regenerated in the tests with the same generator seed/order; store them too (small)
    np.savez_compressed(os.path.join(HERE, "encoder_inputs.npz"), xin=xin.numpy(), xin3=xin3.numpy(),
                        xin_r=xin_r.numpy(), xin40=xin40.numpy())


if __name__ == "__main__":
    main()
