import torch, time
n = 74_181_672 // 8
h_in = torch.zeros(n, dtype=torch.float64).pin_memory()
h_out = torch.zeros(n, dtype=torch.float64).pin_memory()
d_a = torch.zeros(n, dtype=torch.float64, device="cuda")
d_b = torch.zeros(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
def chunked(k=16):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d_a[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_b[i*c:(i+1)*c], non_blocking=True)
mb = n * 8 / 1e6
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both), ("chunked16", chunked)):
    ms = t(fn)
    print(f"{name}: {ms:.3f} ms  ({mb/ms:.1f} GB/s per direction)")

# copies beside kernels: does SM / HBM activity slow the copy engines down?
s3 = torch.cuda.Stream()
A = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
big = torch.zeros(64 * 1024 * 1024, dtype=torch.float64, device="cuda")
def with_fp64():
    with torch.cuda.stream(s3):
        for _ in range(3): torch.mm(A, A)
    chunked()
def with_hbm():
    with torch.cuda.stream(s3):
        for _ in range(6): big.add_(1.0)
    chunked()
def chunked48():
    c = n // 48
    for i in range(48):
        with torch.cuda.stream(s1): d_a[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_b[i*c:(i+1)*c], non_blocking=True)
for name, fn in (("chunked48 (1.5 MB copies)", chunked48), ("chunked16 beside FP64 GEMMs", with_fp64), ("chunked16 beside HBM-bound kernels", with_hbm)):
    ms = t(fn)
    print(f"{name}: {ms:.3f} ms")

# 48 copies per direction spread over 2 / 3 streams per direction: do the per-copy set-up gaps overlap?
def multi(kstreams):
    hs = [torch.cuda.Stream() for _ in range(kstreams)]
    ds = [torch.cuda.Stream() for _ in range(kstreams)]
    c = n // 48
    def fn():
        for i in range(48):
            with torch.cuda.stream(hs[i % kstreams]): d_a[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
            with torch.cuda.stream(ds[i % kstreams]): h_out[i*c:(i+1)*c].copy_(d_b[i*c:(i+1)*c], non_blocking=True)
    return fn
for k in (1, 2, 3, 4):
    print(f"chunked48 over {k} stream(s) per direction: {t(multi(k)):.3f} ms")
