"""Where a Gram-SVD bond step spends its time (wall clock with synchronisation between stages)."""
import sys, time, torch
sys.path.insert(0, '.')
from tensor_networks_b200 import TensorTrain
from tensor_networks_b200 import gramsvd as g
from tensor_networks_b200.utils import delta_svd_dev
d, n = 8, 64
y = None
for j in range(4):
    t = TensorTrain.rand([n] * d, [32] * (d - 1), seed=5001 + j); t.cores[0].mul_(10.0 ** (-3 * j))
    y = t if y is None else y + t
y.clone().gramsvd_round(1e-5)
c = y.cores[3]; r0, nn, r1 = c.shape
m2 = c.reshape(r0 * nn, r1)
def T(f, reps=5):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): out = f()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t) / reps, out
t_gl, gl = T(lambda: g.dev_mm(m2, m2, ta=True))
print("gram GEMM", t_gl)
t_e, (vl, el, _, _) = T(lambda: delta_svd_dev(gl, 0.0))
print("eig via delta_svd", t_e, el.shape)
t_r, _ = T(lambda: g._rounded_sqrt(el))
print("rounded sqrt (host)", t_r)
tmp = g.dev_mm(vl, vl, ta=True)
t_s, _ = T(lambda: delta_svd_dev(tmp, 1e-6))
print("svd tmp", t_s)
t_all, _ = T(lambda: g.gram_eig_and_svd_dev(gl, gl, 1e-6))
print("gram_eig_and_svd_dev", t_all)
