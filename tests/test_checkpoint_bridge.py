"""Checkpoint bridge: a checkpoint written by the REFERENCE code (whole pickled DataParallel model,
SBL/utils.py:22-33) is read WITHOUT the reference on sys.path and loaded into the drop-in modules; weights exported
from the drop-ins load back into the reference modules."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/SBL_Multilingual_Lip_reading"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_checkpoint_round_trip(tmp_path):
    ckpt = tmp_path / "checkpoint.tar"
    # 1. the reference writes a checkpoint exactly like save_checkpoint does (own process: its modules on sys.path)
    writer = textwrap.dedent(f"""
        import sys, torch
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {REF!r})
        from transformer.encoder import Encoder
        from transformer.decoder import Decoder
        from transformer.transformer import Transformer
        from sbl_for_multilingual_lip_reading_b200 import synth
        torch.manual_seed(3)
        enc = Encoder(512, 2, 8, 64, 64, 512, 2048, dropout=0.1, pe_maxlen=5000)
        dec = Decoder(0, 1, 58, 512, 1, 8, 64, 64, 512, 2048, dropout=0.1, tgt_emb_prj_weight_sharing=1, pe_maxlen=5000)
        model = Transformer(enc, dec, None)
        sd = dict(synth.frontend_state_dict(1, prefix="visual_frontend."))
        sd.update(synth.encoder_state_dict(2, 2, prefix="encoder."))
        model.load_state_dict(sd, strict=False)
        state = {{'epoch': 7, 'epochs_since_improvement': 1, 'loss': 0.25,
                 'model': torch.nn.DataParallel(model), 'optimizer': None}}
        torch.save(state, {str(ckpt)!r})
    """)
    subprocess.run([sys.executable, "-c", writer], check=True, timeout=300)
    # 2. read it here: the reference package is NOT importable in this process
    assert REF not in sys.path
    from sbl_for_multilingual_lip_reading_b200 import checkpoint, synth
    from sbl_for_multilingual_lip_reading_b200.encoder import Encoder
    from sbl_for_multilingual_lip_reading_b200.video_frontend import Lipreading
    ck = checkpoint.load_reference_checkpoint(str(ckpt))
    assert ck["epoch"] == 7 and ck["loss"] == 0.25
    sd = ck["state_dict"]
    assert "visual_frontend.frontend3D.0.weight" in sd and "encoder.layer_stack.1.pos_ffn.w_2.bias" in sd
    assert any(k.startswith("decoder.") for k in sd)
    fe, enc = Lipreading(), Encoder(512, 2, 8, 64, 64, 512, 2048)
    rep = checkpoint.load_into_dropins(sd, fe, enc)
    assert rep["frontend"][0] == rep["frontend"][1] == len(fe.state_dict())
    assert rep["encoder"][0] == rep["encoder"][1] == len(enc.state_dict())
    want_f, want_e = synth.frontend_state_dict(1), synth.encoder_state_dict(2, 2)
    for k, v in fe.state_dict().items():
        assert torch.equal(v, want_f[k]), k
    for k, v in enc.state_dict().items():
        assert torch.equal(v, want_e[k]), k
    # 3. a 3-layer encoder only takes what matches in name and shape (the reference's own filter, train.py:98)
    enc3 = Encoder(512, 3, 8, 64, 64, 512, 2048)
    rep3 = checkpoint.load_into_dropins(sd, None, enc3)
    assert rep3["encoder"][0] < rep3["encoder"][1]
    # 4. export -> a `pt` file the reference factory `visual_frontend(pt)` accepts (video_frontend.py:176-190)
    pt = tmp_path / "frontend.pt"
    torch.save(checkpoint.export_reference_state_dict(fe, None, frontend_prefix=""), str(pt))
    reader = textwrap.dedent(f"""
        import sys, torch
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {REF!r})
        from transformer.video_frontend import visual_frontend
        from sbl_for_multilingual_lip_reading_b200 import synth
        m = visual_frontend({str(pt)!r})
        want = synth.frontend_state_dict(1)
        assert all(torch.equal(v, want[k]) for k, v in m.state_dict().items())
        print("reference loaded", len(want), "tensors")
    """)
    out = subprocess.run([sys.executable, "-c", reader], check=True, timeout=300, capture_output=True, text=True).stdout
    assert "reference loaded" in out


def test_stand_in_type_follows_the_pickled_state(tmp_path):
    """A class that cannot be imported becomes an nn.Module stand-in only if its pickled STATE is a module's
    (`_parameters` / `_modules`); the reference's `transformer.optimizer.TransformerOptimizer` (a plain object holding
    an optimizer and counters, SBL/transformer/optimizer.py) stays an attribute bag."""
    pkg = tmp_path / "pkg" / "transformer"
    pkg.mkdir(parents=True)
    (pkg / "optimizer.py").write_text(textwrap.dedent("""
        class TransformerOptimizer(object):
            def __init__(self, k, d_model, warmup_steps):
                self.k, self.init_lr, self.warmup_steps, self.step_num = k, d_model ** (-0.5), warmup_steps, 12
    """))
    (pkg / "tiny.py").write_text(textwrap.dedent("""
        import torch
        class Tiny(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.fc = torch.nn.Linear(3, 2)
    """))
    ckpt = tmp_path / "ck.tar"
    writer = textwrap.dedent(f"""
        import sys, torch
        sys.path.insert(0, {str(tmp_path / 'pkg')!r})
        from transformer.optimizer import TransformerOptimizer
        from transformer.tiny import Tiny
        torch.manual_seed(0)
        torch.save({{'epoch': 1, 'epochs_since_improvement': 0, 'loss': 1.0, 'model': Tiny(),
                    'optimizer': TransformerOptimizer(0.2, 512, 4000)}}, {str(ckpt)!r})
    """)
    subprocess.run([sys.executable, "-c", writer], check=True, timeout=300)
    from sbl_for_multilingual_lip_reading_b200 import checkpoint
    obj = torch.load(str(ckpt), map_location="cpu", pickle_module=checkpoint._tolerant_pickle, weights_only=False)
    assert isinstance(obj["model"], torch.nn.Module) and set(obj["model"].state_dict()) == {"fc.weight", "fc.bias"}
    opt = obj["optimizer"]
    assert not isinstance(opt, torch.nn.Module) and opt.step_num == 12 and opt.warmup_steps == 4000
    assert set(checkpoint.load_reference_checkpoint(str(ckpt))["state_dict"]) == {"fc.weight", "fc.bias"}
