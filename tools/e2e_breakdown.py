"""Where the end-to-end call (numpy in -> numpy out) spends its time beyond the device forward.  Development tool."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import iris.hifigan_pretrained as hp


def t(fn, n=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
    B, T = 16, 862
    torch.manual_seed(0)
    m = hp.HiFiGANModel()
    m.to("cuda:0")
    eng = m.engine
    mel = torch.randn(B, 80, T).numpy()
    pin_in = torch.empty(mel.size, dtype=torch.float32, pin_memory=True).view(B, 80, T)
    out_pin = torch.empty((B, T * eng.hop), dtype=torch.float32, pin_memory=True)
    mel_dev = torch.from_numpy(mel).cuda()
    out_dev = torch.empty(B, T * eng.hop, device="cuda")
    print("copyto pinned   %.3f ms" % t(lambda: np.copyto(pin_in.numpy(), mel, casting="unsafe")))
    print("empty pinned    %.3f ms" % t(lambda: torch.empty((B, T * eng.hop), dtype=torch.float32, pin_memory=True)))
    print("device forward  %.3f ms" % t(lambda: (eng.forward_ptr(mel_dev.data_ptr(), B, T, out_dev.data_ptr(), mode, mel_on_device=True, wave_on_device=True, sync=False), eng.sync())))
    print("pinned forward  %.3f ms" % t(lambda: eng.forward_ptr(pin_in.data_ptr(), B, T, out_pin.data_ptr(), mode)))
    print("engine.forward  %.3f ms (pipelined halves)" % t(lambda: eng.forward(mel, mode)))
    os.environ["HFG_PIPELINE"] = "0"
    print("engine.forward  %.3f ms (HFG_PIPELINE=0: one synchronous call)" % t(lambda: eng.forward(mel, mode)))
    del os.environ["HFG_PIPELINE"]
    h = B // 2
    mel_h = mel_dev[:h].contiguous()
    out_h = torch.empty(h, T * eng.hop, device="cuda")
    print("device fwd B/2  %.3f ms" % t(lambda: (eng.forward_ptr(mel_h.data_ptr(), h, T, out_h.data_ptr(), mode, mel_on_device=True, wave_on_device=True, sync=False), eng.sync())))
    print("2 halves, each synchronous, pinned   %.3f ms" % t(lambda: (eng.forward_ptr(pin_in[:h].data_ptr(), h, T, out_pin[:h].data_ptr(), mode),
                                                                     eng.forward_ptr(pin_in[h:].data_ptr(), B - h, T, out_pin[h:].data_ptr(), mode))))
    print("2 halves, NO_SYNC + one sync, pinned %.3f ms" % t(lambda: (eng.forward_ptr(pin_in[:h].data_ptr(), h, T, out_pin[:h].data_ptr(), mode, sync=False),
                                                                     eng.forward_ptr(pin_in[h:].data_ptr(), B - h, T, out_pin[h:].data_ptr(), mode, sync=False),
                                                                     eng.sync())))
    print("1 whole, NO_SYNC + sync, pinned      %.3f ms" % t(lambda: (eng.forward_ptr(pin_in.data_ptr(), B, T, out_pin.data_ptr(), mode, sync=False), eng.sync())))
    out_np = np.empty((B, T * eng.hop), np.float32)
    print("pageable fwd    %.3f ms" % t(lambda: eng.forward_ptr(mel.ctypes.data, B, T, out_np.ctypes.data, mode)))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        print("d2h pinned 14MB %.3f ms" % t(lambda: out_pin.copy_(out_dev, non_blocking=True)))
        print("h2d pinned 4MB  %.3f ms" % t(lambda: mel_dev.copy_(pin_in, non_blocking=True)))


if __name__ == "__main__":
    main()
