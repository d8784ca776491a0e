import sys, torch, numpy as np
sys.path.insert(0, '.')
from tensor_networks_b200 import TensorTrain, _lib
from tensor_networks_b200 import gramsvd as g
from tensor_networks_b200.tt import workspace, _stream_ptr
d, n = 12, 64
y = None
for j in range(4):
    t = TensorTrain.rand([n] * d, [32] * (d - 1), seed=5001 + j); t.cores[0].mul_(10.0 ** (-3 * j))
    y = t if y is None else y + t
cores = y.cores
last = cores[d - 1].reshape(cores[d - 1].shape[0], -1)
gr = [None] * d
gr[d - 1] = g.dev_mm(last, last, tb=True)
for i in range(d - 2, -1, -1):
    c = cores[i]; r0, nn, r1 = c.shape
    tmp = g.dev_mm(c.reshape(r0 * nn, r1), gr[i + 1]).reshape(r0, nn * r1)
    gr[i] = g.dev_mm(tmp, c.reshape(r0, nn * r1), tb=True)
G = torch.stack(gr[1:]).contiguous()
L = _lib.lib()
count, p = G.shape[0], G.shape[1]
a = torch.empty_like(G); b = torch.empty_like(G); eig = torch.empty((count, p), dtype=torch.float64, device='cuda')
st = torch.zeros((count, 2), dtype=torch.float64, device='cuda')
ws = workspace(L.ttb_gram_eig_batched_workspace_bytes(count, p), G.device, slot="gram_eig")
_lib.check(L.ttb_gram_eig_batched_f64(G.data_ptr(), count, p, a.data_ptr(), b.data_ptr(), eig.data_ptr(), st.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()))
print("sweeps/converged", st.cpu().numpy().tolist())
print("eig range", (eig[:, 0] / eig[:, -1].clamp_min(1e-300)).cpu().numpy())
c = cores[0]; m2 = c.reshape(-1, c.shape[2]); gl = g.dev_mm(m2, m2, ta=True)
G1 = gl[None].contiguous(); a1 = torch.empty_like(G1); b1 = torch.empty_like(G1); e1 = torch.empty((1, G1.shape[1]), dtype=torch.float64, device='cuda'); s1 = torch.zeros((1, 2), dtype=torch.float64, device='cuda')
_lib.check(L.ttb_gram_eig_batched_f64(G1.data_ptr(), 1, G1.shape[1], a1.data_ptr(), b1.data_ptr(), e1.data_ptr(), s1.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()))
print("left gram 0: p", G1.shape[1], "sweeps", s1.cpu().numpy().tolist())
