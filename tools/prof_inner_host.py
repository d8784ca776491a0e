"""Host-side cost of one single-train inner product (enqueue only, no synchronisation) for small trains, split into the
Python pieces (descriptors, workspace query) and the library call."""
import sys, time, torch
sys.path.insert(0, '.')
from tensor_networks_b200 import TensorTrain, _lib
from tensor_networks_b200.tt import workspace, _stream_ptr
L = _lib.lib()
for (d, n, r) in [(20, 4, 4), (20, 20, 40), (20, 20, 80), (20, 20, 160)]:
    a = TensorTrain.rand([n] * d, [r] * (d - 1), seed=1); b = TensorTrain.rand([n] * d, [r] * (d - 1), seed=2)
    for _ in range(5): a.inner_dev(b)
    torch.cuda.synchronize()
    N = 200
    t0 = time.perf_counter()
    for _ in range(N): da, db = a.descriptor(), b.descriptor()
    t_desc = (time.perf_counter() - t0) / N
    t0 = time.perf_counter()
    for _ in range(N): nb = L.ttb_inner_workspace_bytes(da.ref(), db.ref())
    t_wsq = (time.perf_counter() - t0) / N
    ws = workspace(nb, a.device); out = torch.empty((), dtype=torch.float64, device='cuda')
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N): L.ttb_inner_f64(da.ref(), db.ref(), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr())
    t_call = (time.perf_counter() - t0) / N
    torch.cuda.synchronize()
    t_gpu_all = (time.perf_counter() - t0) / N
    t0 = time.perf_counter()
    for _ in range(N): a.inner_dev(b)
    t_full = (time.perf_counter() - t0) / N
    torch.cuda.synchronize()
    t_single = 0.0  # enqueue into an EMPTY stream (pure host cost)
    for _ in range(50):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        L.ttb_inner_f64(da.ref(), db.ref(), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr())
        t_single += (time.perf_counter() - t0) / 50
    torch.cuda.synchronize()
    print(f"d={d} n={n} r={r}: descriptors {1e6*t_desc:.0f} us, workspace query {1e6*t_wsq:.0f} us, library call (enqueue) {1e6*t_call:.0f} us "
          f"(with completion {1e6*t_gpu_all:.0f} us), inner_dev total enqueue {1e6*t_full:.0f} us, library call into an empty stream {1e6*t_single:.0f} us")
