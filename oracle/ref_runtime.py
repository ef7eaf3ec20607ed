"""Run the UNMODIFIED reference modules staged under oracle/_ref (oracle/fetch_ref.py) — TEST INFRASTRUCTURE.

Importers: tests/ (CPU and `-m gpu`), bench.py `--impl reference` / `cpu_baseline`.  Never the product package.

The reference is a flat script bundle whose sub-projects all use the top-level package name `transformer` (and a
top-level `config`), so only one of them can be imported at a time: `load_reference(which)` purges those names from
sys.modules, puts the requested sub-project first on sys.path and imports its hot-path modules.  Instances created
from an earlier load keep working (they hold their classes).
"""
from __future__ import annotations

import contextlib
import importlib
import sys
import types

import torch

from . import fetch_ref

_DIRS = {"sbl": fetch_ref.sbl_dir, "cls": fetch_ref.cls_dir}
_MODS = {"sbl": ("video_frontend", "encoder", "attention", "module", "utils", "transformer", "decoder"),
         "cls": ("video_frontend", "encoder", "attention", "module", "utils", "transformer")}
_loaded = {"which": None}


def available() -> bool:
    return fetch_ref.available() or fetch_ref.fetch(verbose=False)


def ref_dir(which="sbl") -> str:
    return _DIRS[which]()


def load_reference(which="sbl"):
    """-> SimpleNamespace(dir, video_frontend, encoder, ..., Transformer, Encoder, Decoder (sbl only), Lipreading)."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/fetch_ref.py` where /root/reference exists")
    d = ref_dir(which)
    if _loaded["which"] != which:
        for name in [m for m in sys.modules if m == "config" or m == "transformer" or m.startswith("transformer.")]:
            del sys.modules[name]
        for other in _DIRS.values():
            while other() in sys.path:
                sys.path.remove(other())
        sys.path.insert(0, d)
        _loaded["which"] = which
    ns = types.SimpleNamespace(dir=d, which=which)
    for m in _MODS[which]:
        setattr(ns, m, importlib.import_module("transformer." + m))
    ns.Transformer = ns.transformer.Transformer
    ns.Encoder = ns.encoder.Encoder
    ns.Lipreading = ns.video_frontend.Lipreading
    if which == "sbl":
        ns.Decoder = ns.decoder.Decoder
    return ns


@contextlib.contextmanager
def dropout_neutralised(ns):
    """The reference's always-on `F.dropout(x, p=0.5)` (video_frontend.py:122, functional default training=True) made
    an identity for deterministic parity runs (SURVEY.md 8c caveat 1) by swapping the module-level name `F` the call
    resolves through; the reference file itself is untouched."""
    saved = ns.video_frontend.F
    ns.video_frontend.F = types.SimpleNamespace(dropout=lambda x, p=0.5, **kw: x)
    try:
        yield
    finally:
        ns.video_frontend.F = saved


def build_sbl_reference(ns, state_dict=None, n_layers_enc=6, n_layers_dec=6, seed=7):
    """The all-reference SBL model as train.py:58-69 / test.py:86-97 build it (vocab 58, sos 0, eos 1), eval mode, CPU.
    `state_dict`: optional partial {key: tensor} loaded non-strictly (synthetic frontend / encoder weights)."""
    torch.manual_seed(seed)
    enc = ns.Encoder(512, n_layers_enc, 8, 64, 64, 512, 2048, dropout=0.1, pe_maxlen=5000)
    dec = ns.Decoder(0, 1, 58, 512, n_layers_dec, 8, 64, 64, 512, 2048, dropout=0.1, tgt_emb_prj_weight_sharing=1,
                     pe_maxlen=5000)
    model = ns.Transformer(enc, dec, None)
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        assert not unexpected, unexpected
    return model.eval()


def build_cls_reference(ns, state_dict=None, n_layers_enc=3, seed=7):
    """The stage-1 pre-training model (…classify/train.py:37-41): frontend + 3-layer encoder + fc_1500 / fc_2 heads."""
    torch.manual_seed(seed)
    enc = ns.Encoder(512, n_layers_enc, 8, 64, 64, 512, 2048, dropout=0.1, pe_maxlen=5000)
    model = ns.Transformer(enc, None)
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        assert not unexpected, unexpected
    return model


def fp32_exact():
    """Make CUDA fp32 a valid oracle: no TF32 in cuDNN convolutions or cuBLAS matmuls (SURVEY.md 8c caveat 4)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.set_float32_matmul_precision("highest")
    except Exception:
        pass
