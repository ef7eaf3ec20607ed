#!/usr/bin/env python
"""Golden fixture for the input pipeline, produced by the reference's own transforms (build container only):
    python tests/golden/make_golden_input.py
Imports SBL_Multilingual_Lip_reading/cvtransforms.py (ColorNormalize, CenterCrop, RandomCrop) unmodified; `load_file`
(data_gen.py:122-125, `np.load(f) / 255.`) and the frame zero-padding (data_gen.py:291-294) are two-line numpy
statements of a module that cannot be imported here (it pulls in librosa) and are applied literally.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/SBL_Multilingual_Lip_reading")
import cvtransforms  # noqa: E402  (the reference module)

from sbl_for_multilingual_lip_reading_b200 import synth  # noqa: E402


def main():
    u8 = synth.synthetic_u8_clips(1, 29, seed=21)[0].numpy()
    vid = cvtransforms.ColorNormalize(u8 / 255.)
    centre = cvtransforms.CenterCrop(vid, (88, 88))
    vids = np.zeros((30, 88, 88), dtype=np.float32)
    vids[:29] = centre
    random.seed(5)
    rnd = cvtransforms.RandomCrop(vid, (88, 88))
    vids_r = np.zeros((31, 88, 88), dtype=np.float32)
    vids_r[:29] = rnd
    np.savez_compressed(os.path.join(HERE, "input_pipeline.npz"), eval_T30=vids, train_crop_T31=vids_r)
    print("saved", vids.shape, vids_r.shape, float(vids.mean()))


if __name__ == "__main__":
    main()
