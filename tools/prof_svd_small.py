import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from tensor_networks_b200.utils import delta_svd_dev
rng = np.random.default_rng(0)
for p in (48, 64, 96, 128):
    u, _ = np.linalg.qr(rng.standard_normal((p, p))); v, _ = np.linalg.qr(rng.standard_normal((p, p)))
    a = torch.from_numpy((u * 0.8 ** np.arange(p)) @ v.T).cuda()
    for _ in range(2): delta_svd_dev(a, 0.0)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): U, s, svt, info = delta_svd_dev(a, 0.0)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t) / 5 * 1e3
    print(p, "ms", ms)
