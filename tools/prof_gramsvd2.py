"""Accumulated wall time per stage of gramsvd_round on the generic workload (sync after each stage)."""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from tensor_networks_b200 import TensorTrain
from tensor_networks_b200 import gramsvd as g
from tensor_networks_b200.utils import delta_svd_dev
d, n = 12, 64
y = None
for j in range(4):
    t = TensorTrain.rand([n] * d, [32] * (d - 1), seed=5001 + j); t.cores[0].mul_(10.0 ** (-3 * j))
    y = t if y is None else y + t
y.clone().gramsvd_round(1e-5)
acc = {}
def tick(name, f):
    torch.cuda.synchronize(); t = time.perf_counter(); out = f(); torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + 1e3 * (time.perf_counter() - t); return out
tt = y.clone(); cores = tt.cores; eps = 1e-5
last = cores[d - 1].reshape(cores[d - 1].shape[0], -1)
gr = [None] * d
gr[d - 1] = tick("gram_sweep", lambda: g.dev_mm(last, last, tb=True))
for i in range(d - 2, -1, -1):
    c = cores[i]; r0, nn, r1 = c.shape
    tmp = tick("gram_sweep", lambda: g.dev_mm(c.reshape(r0 * nn, r1), gr[i + 1]).reshape(r0, nn * r1))
    gr[i] = tick("gram_sweep", lambda: g.dev_mm(tmp, c.reshape(r0, nn * r1), tb=True))
norm = float(np.sqrt(gr[0].reshape(-1)[0].item())); delta = eps * norm / (d - 1) ** 0.5
ar, br, _ = tick("eig_right_batched", lambda: g.gram_eig_batched_dev(torch.stack(gr[1:])))
for i in range(d - 1):
    c = cores[i]; r0, nn, r1 = c.shape; m2 = c.reshape(r0 * nn, r1)
    gl = tick("gl", lambda: g.dev_mm(m2, m2, ta=True))
    al, bl, _ = tick("eig_l", lambda: g.gram_eig_batched_dev(gl[None]))
    tmp = tick("tmp_gemm", lambda: g.dev_mm(al[0], ar[i], ta=True))
    u, s, svt, _ = tick("svd_mid", lambda: delta_svd_dev(tmp, float(delta)))
    curr = tick("factors", lambda: g.dev_mm(bl[0], u)); nxt = tick("factors", lambda: g.dev_mm(svt, br[i], tb=True))
    rk = curr.shape[1]
    cores[i] = tick("core_update", lambda: g.dev_mm(m2, curr).reshape(r0, nn, rk))
    c1 = cores[i + 1]
    cores[i + 1] = tick("core_update", lambda: g.dev_mm(nxt, c1.reshape(c1.shape[0], -1)).reshape(rk, c1.shape[1], c1.shape[2]))
for k, v in acc.items(): print(f"{k:20s} {v:8.3f} ms total  {v/(d-1):7.3f} per bond")
