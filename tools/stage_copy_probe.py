import numpy as np, torch, time
mel=np.random.randn(16,80,862).astype(np.float32)
dst=torch.empty(mel.size,dtype=torch.float32,pin_memory=True).view(16,80,862)
print("threads", torch.get_num_threads())
for name,fn in (("np.copyto",lambda: np.copyto(dst.numpy(),mel,casting="unsafe")),("torch.copy_",lambda: dst.copy_(torch.from_numpy(mel)))):
    for _ in range(5): fn()
    ts=[]
    for _ in range(50):
        t0=time.perf_counter(); fn(); ts.append(time.perf_counter()-t0)
    ts.sort()
    print(name, f"median {ts[25]*1e3:.3f} ms  min {ts[0]*1e3:.3f}  max {ts[-1]*1e3:.3f}")
