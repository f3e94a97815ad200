"""Per-layer parity of a generator with upsample_initial_channel = 1024 (C = 512 ResBlocks): which layer breaks?"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import torch.nn.functional as F
from iris_tts_b200 import Engine
from iris_tts_b200.engine import GeneratorConfig
from oracle import hifigan_oracle as O

c0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rates = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (8, 8, 2, 2)
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
cfg = GeneratorConfig(80, rates, tuple(2 * r for r in rates), c0, (3, 7, 11), ((1, 3, 5),) * 3)
ocfg = O.OracleConfig(80, cfg.upsample_rates, cfg.upsample_kernel_sizes, c0, cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)
sd = O.random_state_dict(ocfg, seed=1, loud=True)
eng = Engine(cfg, 0); eng.load_state_dict(sd, strict=True); eng.finalize()
w = O.folded_weights(sd)
geo = {n: (kind, cin, cout, k, dil) for n, kind, cin, cout, k, dil, _, _ in O.conv_layers(ocfg)}
names = only or ("conv_pre", "ups.0", "resblocks.0.convs1.0", "resblocks.0.convs2.0", "resblocks.1.convs1.1", "resblocks.2.convs2.2", "ups.1", "resblocks.3.convs1.0", "ups.2", "ups.3", "conv_post")
for name in names:
    kind, cin, cout, k, dil = geo[name]
    torch.manual_seed(5)
    x = torch.randn(2, cin, 150)
    pre = name != "conv_pre"
    xin = F.leaky_relu(x, 0.1) if pre else x
    if kind == "conv":
        ref = F.conv1d(xin.double(), w[name + ".weight"].double(), w[name + ".bias"].double(), dilation=dil, padding=O.get_padding(k, dil)).numpy()
    else:
        u = cfg.upsample_rates[int(name.split(".")[1])]
        ref = F.conv_transpose1d(xin.double(), w[name + ".weight"].double(), w[name + ".bias"].double(), stride=u, padding=(k - u) // 2).numpy()
    if name == "conv_post":
        ref = np.tanh(ref)
    line = []
    for mode in ("fp32", "bf16x3", "bf16"):
        try:
            y = eng.run_layer(name, x.numpy(), pre_lrelu=pre, precision=mode)
            line.append(f"{mode} {np.abs(y - ref).max() / max(1.0, np.abs(ref).max()):.2e}")
        except Exception as exc:  # noqa: BLE001
            line.append(f"{mode} {type(exc).__name__}: {str(exc)[:60]}")
    print(f"{name:24s} {cin:4d}->{cout:4d} k{k} d{dil}  " + "  ".join(line))
