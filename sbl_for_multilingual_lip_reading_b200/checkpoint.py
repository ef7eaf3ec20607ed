"""Checkpoint bridge: read the reference's checkpoints into the drop-in modules and write state
dicts back in the reference's key names — WITHOUT the reference code being importable.

The reference pickles whole objects: `save_checkpoint` stores {'epoch', 'epochs_since_improvement', 'loss', 'model':
nn.DataParallel(Transformer(...)), 'optimizer': TransformerOptimizer} into `checkpoint.tar` / `BEST_checkpoint_*.tar`
(SBL/utils.py:22-33) and `train.py:92-103` / `test.py` resume through `checkpoint['model'].module.state_dict()` with a
"keep what matches in name and shape" filter; `visual_frontend(pt)` loads a plain frontend state dict the same way
(transformer/video_frontend.py:176-190).  Unpickling such a file normally needs `transformer.transformer.Transformer`,
`transformer.optimizer.TransformerOptimizer`, ... on `sys.path`.  Here every class that cannot be imported is replaced
by a stand-in: `nn.Module`-shaped objects only need `_parameters / _buffers / _modules` for `state_dict()`, everything
else becomes an attribute bag.  No arithmetic happens here.

A reference `.tar` is a whole-object pickle and is loaded with `weights_only=False`: like `torch.load` on the reference
side it EXECUTES what the file says — load checkpoints from trusted sources only.
"""
from __future__ import annotations

import pickle
from collections import OrderedDict

import torch
import torch.nn as nn


class _StubModule(nn.Module):
    """Stand-in for a pickled nn.Module subclass whose defining module is not importable."""

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):
        raise RuntimeError("stand-in for a pickled reference module: it only carries parameters")


class _StubObject:
    """Stand-in for any other pickled object (optimizer wrappers, argparse namespaces, ...)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


def _is_module_state(state):
    """A pickled nn.Module carries `_parameters` / `_buffers` / `_modules` in its instance state."""
    return isinstance(state, dict) and "_parameters" in state and "_modules" in state


class _TolerantUnpickler(pickle.Unpickler):
    _stubs = {}

    def find_class(self, module, name):
        try:
            return super().find_class(module, name)   # includes pickle's Python-2 name compatibility mapping
        except (ImportError, AttributeError):          # the defining module / class is not importable here
            key = (module, name)
            cls = self._stubs.get(key)
            if cls is None:
                # The pickled INSTANCE STATE decides what the stand-in is (BUILD time, `__setstate__` below): state with
                # `_parameters` / `_modules` makes it an nn.Module stand-in, anything else (the reference's
                # `transformer.optimizer.TransformerOptimizer`, argparse namespaces, ...) stays an attribute bag.
                def __setstate__(self, state, _name=name, _module=module):
                    if _is_module_state(state):
                        self.__class__ = _module_stub_class(_module, _name)
                        nn.Module.__init__(self)
                        nn.Module.__setstate__(self, state)
                    else:
                        _StubObject.__setstate__(self, state)
                cls = type(name, (_StubObject,), {"__module__": module, "_sblk_stub": True, "__setstate__": __setstate__})
                self._stubs[key] = cls
            return cls


_MODULE_STUBS = {}


def _module_stub_class(module, name):
    key = (module, name)
    cls = _MODULE_STUBS.get(key)
    if cls is None:
        cls = _MODULE_STUBS[key] = type(name, (_StubModule,), {"__module__": module, "_sblk_stub": True})
    return cls


class _tolerant_pickle:
    """`pickle_module` for torch.load: the standard pickle with the tolerant class lookup."""
    __name__ = "pickle"
    Unpickler = _TolerantUnpickler
    load = staticmethod(lambda f, **kw: _TolerantUnpickler(f, **kw).load())
    loads = staticmethod(pickle.loads)
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
    DEFAULT_PROTOCOL = pickle.DEFAULT_PROTOCOL
    PickleError = pickle.PickleError
    PicklingError = pickle.PicklingError
    UnpicklingError = pickle.UnpicklingError


def load_reference_checkpoint(path, map_location="cpu"):
    """-> {'state_dict': OrderedDict (reference key names, `module.` of DataParallel stripped), 'epoch', 'loss',
    'epochs_since_improvement'} from a reference `checkpoint.tar`, a bare pickled model, or a plain state dict (`pt`)."""
    obj = torch.load(path, map_location=map_location, pickle_module=_tolerant_pickle, weights_only=False)
    meta = {}
    if isinstance(obj, dict) and "model" in obj and not torch.is_tensor(obj["model"]):
        meta = {k: obj.get(k) for k in ("epoch", "epochs_since_improvement", "loss")}
        obj = obj["model"]
    if isinstance(obj, nn.Module):
        if hasattr(obj, "module") and isinstance(getattr(obj, "module"), nn.Module):   # nn.DataParallel wrapper
            obj = obj.module
        sd = obj.state_dict()
    elif isinstance(obj, dict):
        sd = obj
    else:
        raise RuntimeError(f"{path}: neither a reference checkpoint, a pickled model nor a state dict ({type(obj)})")
    out = OrderedDict()
    for k, v in sd.items():
        if torch.is_tensor(v):
            out[k[len("module."):] if k.startswith("module.") else k] = v
    meta["state_dict"] = out
    return meta


def _matching(sd, prefix, model):
    own = model.state_dict()
    picked = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    # the reference's filter (train.py:98, video_frontend.py:184): keep entries that match in name and shape
    return {k: v for k, v in picked.items() if k in own and tuple(v.shape) == tuple(own[k].shape)}, own


def load_into_dropins(state_dict, frontend=None, encoder=None, frontend_prefix=None, encoder_prefix=None):
    """Load the visual-frontend / encoder entries of a reference state dict (SBL: `visual_frontend.*`, `encoder.*`;
    stage-1 classifier: `visual_frontend.*`, `encoder_v.*`; a bare `pt` file: no prefix) into the drop-in modules with
    the reference's own name+shape filter.  Returns {'frontend': (loaded, total), 'encoder': (loaded, total)}."""
    report = {}
    keys = list(state_dict.keys())

    def pick_prefix(cands):
        for c in cands:
            if any(k.startswith(c) for k in keys):
                return c
        return ""

    if frontend is not None:
        pre = frontend_prefix if frontend_prefix is not None else pick_prefix(("visual_frontend.", "lipreading."))
        picked, own = _matching(state_dict, pre, frontend)
        own.update(picked)
        frontend.load_state_dict(own)
        report["frontend"] = (len(picked), len(own))
    if encoder is not None:
        pre = encoder_prefix if encoder_prefix is not None else pick_prefix(("encoder.", "encoder_v."))
        picked, own = _matching(state_dict, pre, encoder)
        own.update(picked)
        encoder.load_state_dict(own)
        report["encoder"] = (len(picked), len(own))
    return report


def export_reference_state_dict(frontend=None, encoder=None, frontend_prefix="visual_frontend.",
                                encoder_prefix="encoder."):
    """State dict of the drop-ins under the reference `Transformer`'s key names (fp32 CPU tensors), ready for
    `reference_model.load_state_dict(..., strict=False)` or for `torch.save` as a `pt` file."""
    out = OrderedDict()
    if frontend is not None:
        for k, v in frontend.state_dict().items():
            out[frontend_prefix + k] = v.detach().cpu()
    if encoder is not None:
        for k, v in encoder.state_dict().items():
            out[encoder_prefix + k] = v.detach().cpu()
    return out
