"""Golden vectors for the f3 oracle (oracle/heads_oracle.py) from the reference's OWN TemporalAttention class
(DG-SCT/AVE/nets/net_trans.py:183-249), imported read-only from /root/reference with the stub recipe of SURVEY.md
appendix A.  Forward hooks capture the encoder / decoder outputs that feed the gate tail; the module's three return values
and the gradients w.r.t. those captured tensors and the gate parameters are stored (fp64).
Run once in the build container:  python tests/golden/make_golden_f3.py"""
import importlib
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch
import transformers.activations  # noqa: F401  (import the REAL transformers before stubbing timm)

HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    sys.modules[name] = m
    return m


def main():
    for n in ["ipdb", "timm", "timm.models", "timm.models.vision_transformer", "timm.models.layers", "loralib", "torchlibrosa",
              "torchlibrosa.stft", "torchlibrosa.augmentation", "h5py", "librosa"]:
        _stub(n)
    sys.modules["ipdb"].set_trace = lambda *a, **k: None
    sys.modules["timm.models.vision_transformer"].Attention = object
    sys.modules["timm.models.layers"].to_2tuple = lambda x: (x, x)
    sys.modules["timm.models.layers"].trunc_normal_ = lambda *a, **k: None
    for n in ("Spectrogram", "LogmelFilterBank"):
        setattr(sys.modules["torchlibrosa.stft"], n, object)
    sys.modules["torchlibrosa.augmentation"].SpecAugmentation = object
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference/DG-SCT/AVE")
    net = importlib.import_module("nets.net_trans")
    torch.manual_seed(7)
    ta = net.TemporalAttention().double().eval()          # eval: dropout off, deterministic
    cap = {}

    def grab(name):
        def hook(mod, inp, out):
            out.retain_grad()
            cap[name] = out
        return hook

    ta.video_encoder.register_forward_hook(grab("v_kv"))
    ta.audio_encoder.register_forward_hook(grab("a_kv"))
    ta.video_decoder.register_forward_hook(grab("vq"))
    ta.audio_decoder.register_forward_hook(grab("aq"))
    B, T = 4, 10
    f_v = torch.randn(B, T, 1536, dtype=torch.float64)
    f_a = torch.randn(B, T, 768, dtype=torch.float64)
    vq_out, aq_out, av = ta(f_v, f_a)
    R = [torch.randn_like(t) for t in (vq_out, aq_out, av)]
    (vq_out * R[0]).sum().add((aq_out * R[1]).sum()).add((av * R[2]).sum()).backward()
    out = {k: v.detach().numpy() for k, v in cap.items()}
    out.update({"grad_" + k: v.grad.numpy() for k, v in cap.items()})
    out.update(vq_out=vq_out.detach().numpy(), aq_out=aq_out.detach().numpy(), av=av.detach().numpy(),
               R_vq=R[0].numpy(), R_aq=R[1].numpy(), R_av=R[2].numpy(), gamma=np.float64(ta.gamma),
               wa=ta.audio_gated[0].weight.detach().numpy(), ba=ta.audio_gated[0].bias.detach().numpy(),
               wv=ta.video_gated[0].weight.detach().numpy(), bv=ta.video_gated[0].bias.detach().numpy(),
               grad_wa=ta.audio_gated[0].weight.grad.numpy(), grad_ba=ta.audio_gated[0].bias.grad.numpy(),
               grad_wv=ta.video_gated[0].weight.grad.numpy(), grad_bv=ta.video_gated[0].bias.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "f3_temporal_gates.npz"), **{k: np.asarray(v, dtype=np.float32) if False else v for k, v in out.items()})
    print("wrote f3_temporal_gates.npz:", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
