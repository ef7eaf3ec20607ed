"""CPU restatement (numpy) of the reference input pipeline for one video clip — TEST INFRASTRUCTURE ONLY (imported by
tests/ and bench.py's checks, never by the product package).

Follows SBL_Multilingual_Lip_reading/data_gen.py:122-125 (`load_file`: np.load(...) / 255.), cvtransforms.py:44-48
(`ColorNormalize`), cvtransforms.py:7-20 (`CenterCrop`), cvtransforms.py:23-33 (`RandomCrop`, offsets passed in) and
data_gen.py:291-294 (zero-padding of the clip to a fixed frame count, in normalised space, float32).
Pinned by tests/golden/input_pipeline.npz, produced by the reference's own cvtransforms functions
(tests/golden/make_golden_input.py).
"""
import numpy as np

MEAN, STD = 0.413621, 0.1700239


def load_and_normalize(u8):
    """uint8 [T,H,W] -> float64 [T,H,W]: `arrays / 255.` then `(batch_img - mean) / std`."""
    arrays = np.asarray(u8) / 255.
    return (arrays - MEAN) / STD


def center_crop(batch_img, size=(88, 88)):
    """cvtransforms.CenterCrop: x1 = int(round((w - tw)) / 2.), y1 = int(round((h - th)) / 2.)."""
    w, h = batch_img[0].shape[1], batch_img[0].shape[0]
    th, tw = size
    x1 = int(round((w - tw)) / 2.)
    y1 = int(round((h - th)) / 2.)
    return np.stack([f[y1:y1 + th, x1:x1 + tw] for f in batch_img]), (y1, x1)


def crop_at(batch_img, offsets_yx, size=(88, 88)):
    """cvtransforms.RandomCrop with the per-frame (y1, x1) it drew (x1 = randint(0, 8) then y1 = randint(0, 8))."""
    th, tw = size
    return np.stack([f[y1:y1 + th, x1:x1 + tw] for f, (y1, x1) in zip(batch_img, offsets_yx)])


def pad_frames(vid, t_out):
    """data_gen.py:291-294: vids = np.zeros((t_out, w, h), float32); vids[:length] = vid."""
    length, h, w = vid.shape
    vids = np.zeros((t_out, h, w), dtype=np.float32)
    vids[:length] = vid
    return vids


def eval_clip(u8, t_out):
    """Test-split path of AiShellDataset.__getitem__ for an LRW clip: load, normalise, centre-crop, pad."""
    vid, _ = center_crop(load_and_normalize(u8))
    return pad_frames(vid, t_out)
