"""dgsct_b200.patch() -- the actual drop-in: import a reference task module (stubbing its absent third-party imports as in
SURVEY.md appendix A), swap ``VisualAdapter`` for the engine class, and check that the task network's own construction code
(the argument lists of DG-SCT/AVE/nets/net_trans.py:807-845) builds engine adapters with the reference's state_dict.
Needs /root/reference (present in the build container, absent on the GPU box: skipped there)."""
import importlib
import importlib.machinery
import os
import subprocess
import sys
import textwrap

import pytest

REF = "/root/reference/DG-SCT/AVE"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import sys, types, importlib, importlib.machinery
    import torch, transformers.activations
    def _stub(name):
        m = types.ModuleType(name); m.__spec__ = importlib.machinery.ModuleSpec(name, None); m.__path__ = []
        sys.modules[name] = m; return m
    for n in ["ipdb", "timm", "timm.models", "timm.models.vision_transformer", "timm.models.layers", "loralib",
              "torchlibrosa", "torchlibrosa.stft", "torchlibrosa.augmentation", "h5py", "librosa"]:
        _stub(n)
    sys.modules["ipdb"].set_trace = lambda *a, **k: None
    sys.modules["timm.models.vision_transformer"].Attention = object
    sys.modules["timm.models.layers"].to_2tuple = lambda x: (x, x)
    sys.modules["timm.models.layers"].trunc_normal_ = lambda *a, **k: None
    for n in ("Spectrogram", "LogmelFilterBank"): setattr(sys.modules["torchlibrosa.stft"], n, object)
    sys.modules["torchlibrosa.augmentation"].SpecAugmentation = object
    sys.dont_write_bytecode = True
    sys.path.insert(0, "%(ref)s")
    sys.path.insert(0, "%(root)s")
    net = importlib.import_module("nets.net_trans")
    RefAdapter = net.VisualAdapter
    import dgsct_b200
    Bound = dgsct_b200.patch(net, "ave")
    assert net.VisualAdapter is Bound and Bound.__name__ == "VisualAdapter" and issubclass(Bound, dgsct_b200.VisualAdapter)
    opt = types.SimpleNamespace(is_multimodal=1, num_conv_group=2, is_before_layernorm=1, is_post_layernorm=1)
    # the exact keyword list of the reference's instantiation sites (net_trans.py:807-845)
    kw = dict(input_dim=768, output_dim=768, adapter_kind="bottleneck", dim_list=[1, 2], layer_idx=3, reduction_factor=8, opt=opt,
              use_bn=True, use_gate=True, num_tk=32, conv_dim_in=256, conv_dim_out=144, linear_in=384, linear_out=768)
    eng, ref = net.VisualAdapter(**kw), RefAdapter(**kw)
    assert eng.variant == "ave"
    sd_e, sd_r = eng.state_dict(), ref.state_dict()
    assert list(sd_e) == list(sd_r), (sorted(set(sd_e) ^ set(sd_r)))
    assert all(sd_e[k].shape == sd_r[k].shape and sd_e[k].dtype == sd_r[k].dtype for k in sd_r)
    eng.load_state_dict(sd_r)                      # a reference checkpoint loads strictly
    ref.load_state_dict(eng.state_dict())          # and an engine checkpoint loads into the reference class
    # same trainable set as the reference's name-based freeze policy sees (main_trans.py:211-270)
    assert [n for n, _ in eng.named_parameters()] == [n for n, _ in ref.named_parameters()]
    # the engine has no CPU path: calling it on CPU tensors must raise, not fall back
    x = torch.randn(2, 768, 144, 1); o = torch.randn(2, 384, 256, 1)
    try:
        eng.eval()(x, o)
        raise SystemExit("engine ran on CPU tensors")
    except RuntimeError as e:
        assert "CUDA" in str(e)
    print("PATCH-OK")
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_patch_swaps_the_class_inside_the_reference_module():
    out = subprocess.run([sys.executable, "-c", SCRIPT % dict(ref=REF, root=ROOT)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "PATCH-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
