#!/usr/bin/env python
"""Top-1 labels of 1,000 synthetic clips computed by THE REFERENCE ITSELF (BASELINE.json north_star:
"identical top-1 class ... on 1,000 synthetic clips").

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_top1.py

The unmodified reference `Lipreading` (always-on dropout avoided through `_frontend_forward`, as in
make_golden.py) and 6-layer `Encoder` from /root/reference/SBL_Multilingual_Lip_reading run on CPU in fp32 with the
seeded synthetic weights of `synth`; the stage-1 heads fc_1500 / fc_2 (…classify/transformer/transformer.py:13-14)
are applied by `synth.classify`.  Stored per clip: top-1 word class, top-1 minus top-2 logit margin, language
top-1 and margin.  Clips: chunk c (40 clips x 29 frames) = synth.structured_clips(40, 29, seed=5000 + c).

With random weights the encoder output is dominated by clip-independent terms (SURVEY.md §8c caveat 3): every clip
lands in the same class.  A second, discriminative labelling is therefore stored too: the word head applied to the
pooled output CENTRED by its mean over the 1,000 reference outputs (`mu`, stored), which spreads the clips over the
classes and makes the comparison sensitive to the encoder's numerical error.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/SBL_Multilingual_Lip_reading"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from transformer.encoder import Encoder as RefEncoder  # noqa: E402
from transformer.video_frontend import Lipreading as RefLipreading  # noqa: E402

from sbl_for_multilingual_lip_reading_b200 import synth  # noqa: E402

CHUNK, CHUNKS, T = 40, 25, 29


def main():
    torch.set_grad_enabled(False)
    fe = RefLipreading()
    fe.load_state_dict(synth.frontend_state_dict(1))
    enc = RefEncoder(512, 6, 8, 64, 64, 512, 2048, dropout=0.1, pe_maxlen=5000)
    enc.load_state_dict(synth.encoder_state_dict(2, 6))
    fe, enc = fe.eval(), enc.eval()
    heads = synth.classifier_heads(9)
    top1, margin, lang, lang_margin, pooled = [], [], [], [], []
    t0 = time.time()
    for c in range(CHUNKS):
        x = synth.structured_clips(CHUNK, T, seed=5000 + c)
        feat = fe._frontend_forward(x).view(CHUNK, T, 512)
        out, = enc(feat, [T] * CHUNK)
        logits, ll = synth.classify(out, heads)
        pooled.append(out.mean(dim=1))
        v, i = logits.topk(2, dim=1)
        top1.append(i[:, 0]); margin.append(v[:, 0] - v[:, 1])
        v2, i2 = ll.topk(2, dim=1)
        lang.append(i2[:, 0]); lang_margin.append(v2[:, 0] - v2[:, 1])
        print(f"chunk {c + 1}/{CHUNKS}  {time.time() - t0:.0f}s", flush=True)
    pooled = torch.cat(pooled)
    mu = pooled.mean(dim=0)
    lc = (pooled - mu) @ heads["fc_1500.weight"].t()
    vc, ic = lc.topk(2, dim=1)
    print("distinct classes: plain", len(set(torch.cat(top1).tolist())), "centred", len(set(ic[:, 0].tolist())))
    np.savez_compressed(os.path.join(HERE, "top1_1000.npz"), mu=mu.numpy(),
                        top1_centred=ic[:, 0].numpy().astype(np.int16), margin_centred=(vc[:, 0] - vc[:, 1]).numpy(),
                        top1=torch.cat(top1).numpy().astype(np.int16), margin=torch.cat(margin).numpy(),
                        lang=torch.cat(lang).numpy().astype(np.int8), lang_margin=torch.cat(lang_margin).numpy(),
                        chunk=np.int32(CHUNK), chunks=np.int32(CHUNKS), frames=np.int32(T))


if __name__ == "__main__":
    main()
