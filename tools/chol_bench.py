import ctypes, torch
torch.zeros(1, device="cuda")
L = ctypes.CDLL("tensor_networks_b200/libttb200.so")
L.ttb_debug_chol_bench_us.restype = ctypes.c_double
print(64, L.ttb_debug_chol_bench_us(64, 20))
