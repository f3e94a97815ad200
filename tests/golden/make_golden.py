"""Generate tests/golden/*.npz by running the REFERENCE's own module.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

It loads ``/root/reference/src/iris/hifigan_pretrained.py`` by file path (under
the alias ``_ref_hifigan`` so it cannot clash with this repo's ``iris``
package), builds ``HiFiGANModel`` under ``torch.manual_seed(0)`` exactly as
SURVEY.md section 8(c) standardises, and stores for each case: the mel, the
reference waveform, strided samples of intermediate activations (forward
hooks on the reference's own submodules) and per-tensor weight checksums, so
that the oracle's seeded re-creation of the weights can be verified anywhere.
Nothing here is shipped or imported by the product.
"""
import importlib.util
import json
import os
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/iris/hifigan_pretrained.py"
TAP_STRIDE = 97

CASES = {
    # name: (ctor kwargs, weight seed, loud, mel seed, B, T, realistic)
    "v1_default": ({}, 0, False, 1234, 2, 24, False),
    "v1_loud": ({}, 0, True, 1234, 2, 24, False),
    "v1_realistic_odd": ({}, 0, True, 77, 1, 17, True),
    "v2_loud": ({"upsample_initial_channel": 128}, 0, True, 1234, 2, 20, False),
    "v3_loud": (
        {
            "upsample_rates": [8, 8, 4],
            "upsample_kernel_sizes": [16, 16, 8],
            "upsample_initial_channel": 256,
            "resblock_kernel_sizes": [3, 5, 7],
            "resblock_dilation_sizes": [[1, 2], [2, 6], [3, 12]],
        },
        0, True, 1234, 2, 20, False,
    ),
}


def load_reference():
    spec = importlib.util.spec_from_file_location("_ref_hifigan", REF)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_ref_hifigan"] = mod
    spec.loader.exec_module(mod)
    return mod


def make_loud(model):
    torch.manual_seed(1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("weight_g"):
                p.mul_(torch.empty_like(p).uniform_(1.0, 3.0))


def main():
    warnings.filterwarnings("ignore")
    ref = load_reference()
    for case, (kw, wseed, loud, mseed, B, T, realistic) in CASES.items():
        torch.manual_seed(wseed)
        model = ref.HiFiGANModel(**kw).eval()
        if loud:
            make_loud(model)
        torch.manual_seed(mseed)
        mel = torch.randn(B, 80, T)
        if realistic:
            mel = mel * 2.0 - 5.0

        taps = {}
        hooks = []

        def hook(name):
            def f(_m, _i, o):
                taps[name] = o.detach().clone()
            return f

        hooks.append(model.conv_pre.register_forward_hook(hook("conv_pre")))
        for i, u in enumerate(model.ups):
            hooks.append(u.register_forward_hook(hook(f"ups.{i}")))
        for n, rb in enumerate(model.resblocks):
            hooks.append(rb.register_forward_hook(hook(f"resblocks.{n}")))
        hooks.append(model.conv_post.register_forward_hook(hook("conv_post")))
        with torch.no_grad():
            out = model(mel)
        for h in hooks:
            h.remove()

        data = {"mel": mel.numpy(), "out": out.numpy()}
        for k, v in taps.items():
            flat = v.reshape(-1)
            data["tap:" + k] = flat[::TAP_STRIDE].numpy().copy()
            data["tapstat:" + k] = np.array([float(v.double().mean()), float(v.double().std()), float(v.abs().max())])
        sums = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in model.state_dict().items()}
        data["weights_json"] = np.frombuffer(json.dumps(sums).encode(), dtype=np.uint8)
        meta = {"kwargs": kw, "weight_seed": wseed, "loud": loud, "mel_seed": mseed, "B": B, "T": T,
                "realistic": realistic, "tap_stride": TAP_STRIDE, "torch": torch.__version__}
        data["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, case + ".npz"), **data)
        print(case, "out", tuple(out.shape), "mean %.5f std %.5f max|.| %.5f" % (out.mean(), out.std(), out.abs().max()))

    # Public-API case: infer_hifigan through a checkpoint file (hifigan_pretrained.py:286-317)
    torch.manual_seed(0)
    model = ref.HiFiGANModel().eval()
    make_loud(model)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "generator.ckpt")
        torch.save(model.state_dict(), p)
        torch.manual_seed(1234)
        mel = torch.randn(1, 80, 12).numpy()
        a3 = ref.infer_hifigan(mel, checkpoint_path=p)            # [1,80,T] -> [N]
        a2 = ref.infer_hifigan(mel[0], checkpoint_path=p)         # [80,T]   -> [N]
        gen = ref.get_pretrained_hifigan(p)
        g3 = gen(mel)                                             # [1,N]
        g64 = gen(mel.astype(np.float64))
        np.savez_compressed(os.path.join(HERE, "api_shapes.npz"), mel=mel, infer3=a3, infer2=a2, call3=g3, call64=g64)
        print("api", a3.shape, a2.shape, g3.shape, g64.dtype)


if __name__ == "__main__":
    main()
