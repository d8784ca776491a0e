"""Times the DMMA GEMM on the environment-step shapes for every tile / split-K choice
(GPU box only).  Prints one JSON line per configuration; used to tune the heuristic."""

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensor_networks_b200 import _lib  # noqa: E402
from tensor_networks_b200.tt import workspace  # noqa: E402

L = _lib.lib()
TILES = {0: "128x128", 1: "128x112", 2: "64x64", 3: "128x64", 4: "128x128w16"}


def time_gemm(M, N, K, a_kc, b_kc, tile, splits, reps=20, nbuf=6):
    # rotate over nbuf operand sets so that inputs do not sit in L2 between repetitions
    As = [torch.randn((M, K) if a_kc else (K, M), dtype=torch.float64, device="cuda") for _ in range(nbuf)]
    Bs = [torch.randn((N, K) if b_kc else (K, N), dtype=torch.float64, device="cuda") for _ in range(nbuf)]
    C = torch.empty((M, N), dtype=torch.float64, device="cuda")
    ws = workspace(max(L.ttb_gemm_workspace_bytes(M, N, K), 64 * M * N * 8 if splits != 1 else 0), C.device, "sweep")

    def run(i):
        A, B = As[i % nbuf], Bs[i % nbuf]
        sAm, sAk = (A.stride(0), 1) if a_kc else (1, A.stride(0))
        sBk, sBn = (1, B.stride(0)) if b_kc else (B.stride(0), 1)
        _lib.check(L.ttb_gemm_f64_ex(M, N, K, 1.0, A.data_ptr(), sAm, sAk, B.data_ptr(), sBk, sBn, 0.0,
                                     C.data_ptr(), N, tile, splits, ws.data_ptr(), ws.numel(), None))

    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12


def cublas(M, N, K, reps=20):
    A = torch.randn(M, K, dtype=torch.float64, device="cuda")
    B = torch.randn(K, N, dtype=torch.float64, device="cuda")
    for _ in range(3):
        A @ B
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        A @ B
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12


if __name__ == "__main__":
    cases = [
        ("gemm1 T=E.B", 256, 8192, 256, True, False, [(0, 1), (1, 1), (2, 1), (3, 1), (4, 1), (-1, 0)]),
        ("gemm2 E'=A^T.T", 256, 256, 8192, False, False,
         [(0, 37), (4, 37), (2, 18), (2, 9), (3, 18), (3, 36), (-1, 0)]),
        ("push core.R^T", 16384, 256, 256, True, True, [(0, 1), (4, 1), (3, 1), (-1, 0)]),
        ("square 4096", 4096, 4096, 4096, True, False, [(0, 1), (4, 1), (3, 1), (2, 1)]),
        ("proj C=P.Q^T", 32, 224, 16384, True, True, [(2, 16), (2, 64), (-1, 0)]),
        ("proj P-=C.Q", 32, 16384, 224, True, False, [(2, 1), (3, 1), (-1, 0)]),
    ]
    for name, M, N, K, akc, bkc, cfgs in cases:
        ms, tf = cublas(M, N, K)
        print(json.dumps({"case": name, "impl": "cublas", "ms": round(ms, 4), "tflops": round(tf, 2)}), flush=True)
        for tile, splits in cfgs:
            try:
                ms, tf = time_gemm(M, N, K, akc, bkc, tile, splits)
                print(json.dumps({"case": name, "tile": TILES.get(tile, "auto"), "splits": splits,
                                  "ms": round(ms, 4), "tflops": round(tf, 2)}), flush=True)
            except Exception as exc:
                print(json.dumps({"case": name, "tile": tile, "splits": splits, "error": repr(exc)}), flush=True)
