"""Library comparators (sanity ceilings only, never on the product path): cuSOLVER via torch."""
import torch, time
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
a = torch.randn(16384, 256, dtype=torch.float64, device="cuda")
b = torch.randn(8192, 256, dtype=torch.float64, device="cuda")
s = torch.randn(256, 256, dtype=torch.float64, device="cuda")
print("torch.linalg.qr 16384x256 ms", t(lambda: torch.linalg.qr(a)))
print("torch.linalg.qr 8192x256 ms", t(lambda: torch.linalg.qr(b)))
print("torch.linalg.svd 256x256 ms", t(lambda: torch.linalg.svd(s)))
for drv in ("gesvdj", "gesvd", "gesvda"):
    try:
        print("svd", drv, "256x256 ms", t(lambda: torch.linalg.svd(s, driver=drv)))
    except Exception as e:
        print(drv, "failed", e)
print("torch.linalg.eigh 256x256 ms", t(lambda: torch.linalg.eigh(s @ s.T)))
print("gram 256x16384x256 ms", t(lambda: a.T @ a))
print("chol 256 ms", t(lambda: torch.linalg.cholesky(a.T @ a)))
