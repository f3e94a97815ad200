"""Generate tests/golden/pickled_model.ckpt: a MODEL OBJECT pickled by the REFERENCE module.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_pickled_model.py

``HiFiGANGenerator.__init__`` accepts a checkpoint that holds a whole pickled module and uses it as is
(/root/reference/src/iris/hifigan_pretrained.py:168-171).  Such a pickle names the class
``iris.hifigan_pretrained.HiFiGANModel``, so the reference module is loaded here under exactly that
name; a small NON-default architecture keeps the fixture small (and proves the drop-in reads the
architecture from the pickled submodules, not from defaults).  The expected state dict checksums,
a mel and the reference's own output are stored beside it.
"""
import importlib.util
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/iris/hifigan_pretrained.py"
KW = dict(in_channels=80, upsample_rates=[4, 4], upsample_kernel_sizes=[8, 8], upsample_initial_channel=32,
          resblock_kernel_sizes=[3, 5], resblock_dilation_sizes=[[1, 2], [2, 6]])


def main():
    warnings.filterwarnings("ignore")
    pkg = types.ModuleType("iris")
    pkg.__path__ = []
    spec = importlib.util.spec_from_file_location("iris.hifigan_pretrained", REF)
    ref = importlib.util.module_from_spec(spec)
    sys.modules["iris"] = pkg
    sys.modules["iris.hifigan_pretrained"] = ref
    spec.loader.exec_module(ref)
    torch.manual_seed(3)
    model = ref.HiFiGANModel(**KW).eval()
    torch.manual_seed(1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("weight_g"):
                p.mul_(torch.empty_like(p).uniform_(1.0, 3.0))
    torch.manual_seed(5)
    mel = torch.randn(2, 80, 9)
    with torch.no_grad():
        out = model(mel)
    torch.save(model, os.path.join(HERE, "pickled_model.ckpt"))
    sums = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, "pickled_model_expected.npz"), mel=mel.numpy(), out=out.numpy(),
                        weights_json=np.frombuffer(json.dumps(sums).encode(), dtype=np.uint8),
                        kwargs_json=np.frombuffer(json.dumps(KW).encode(), dtype=np.uint8))
    print("pickled model:", os.path.getsize(os.path.join(HERE, "pickled_model.ckpt")), "bytes; out", tuple(out.shape), float(out.std()))


if __name__ == "__main__":
    main()
