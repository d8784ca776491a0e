"""GPU-box experiment: how fast is cudaHostRegister / Unregister on pageable numpy memory (per chunk size, threads)?
Decides whether pinning the caller's buffers on the fly can replace the staging memcpy of csrc/staging.cu."""
import ctypes, time, threading, numpy as np, torch
torch.cuda.init()
rt = ctypes.CDLL("libcudart.so.12") if False else None
import os, glob
cand = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = ctypes.CDLL(cand[0])
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
rt.cudaHostUnregister.argtypes = [ctypes.c_void_p]
N = 1 << 30  # 1 GiB
a = np.random.rand(N // 8)
base = a.ctypes.data
print("cudart", cand[0])
for chunk_mb in (4, 32, 256, 1024):
    chunk = chunk_mb << 20
    n = N // chunk
    for nthreads in (1, 4, 8):
        def work(tid, what):
            for i in range(tid, n, nthreads):
                p = base + i * chunk
                if what == 0:
                    rc = rt.cudaHostRegister(ctypes.c_void_p(p), chunk, 0)
                else:
                    rc = rt.cudaHostUnregister(ctypes.c_void_p(p))
                assert rc == 0, rc
        out = []
        for what in (0, 1):
            ths = [threading.Thread(target=work, args=(t, what)) for t in range(nthreads)]
            t0 = time.perf_counter()
            for t in ths: t.start()
            for t in ths: t.join()
            out.append(time.perf_counter() - t0)
        print(f"chunk {chunk_mb:5d} MB threads {nthreads}: register {N / out[0] / 1e9:6.1f} GB/s, unregister {N / out[1] / 1e9:6.1f} GB/s")
# second registration of the same (already touched) memory
