// Tensor-train descriptor shared by the host-side drivers.
#pragma once

#include "common.cuh"

namespace ttb {

// A tensor train with d cores.  n[k] = mode size, r[k] = bond rank to the left
// of core k (r[0] = r[d] = 1).  core[k] points to DEVICE memory holding the
// C-order array (r[k], n[k], r[k+1]) -- byte-identical to the reference's cores
// (pytens/algs.py:1188-1216) with the unit bonds of the first/last core made
// explicit.  n, r and core themselves are HOST arrays.
struct TTDesc {
    int d;
    const int64_t* n;
    const int64_t* r;
    double* const* core;
};

int validate(const TTDesc& t, const char* what);

// --- inner product (inner.cu) ---
size_t inner_workspace_bytes(const TTDesc& a, const TTDesc& b);
int inner(const TTDesc& a, const TTDesc& b, double* out_dev, void* ws, size_t ws_bytes,
          cudaStream_t stream);

// persistent fused sweep (inner_fused.cu); kUnsupported when the shapes do not qualify
size_t inner_fused_workspace_bytes(const TTDesc& a, const TTDesc& b);
int inner_fused(const TTDesc& a, const TTDesc& b, double* out_dev, void* ws, size_t ws_bytes,
                cudaStream_t stream, const int* ready_dev = nullptr, int* fail_dev = nullptr);

// TMA-staged persistent sweep for bond ranks <= 256 (inner_tma.cu); kUnsupported when the shapes do not qualify
size_t inner_tma_workspace_bytes(const TTDesc& a, const TTDesc& b);
int inner_tma(const TTDesc& a, const TTDesc& b, double* out_dev, void* ws, size_t ws_bytes, cudaStream_t stream,
              const int* ready_dev = nullptr, int* fail_dev = nullptr);

// <A, B> of two trains whose cores still sit in (pinned) HOST memory: the cores are copied to the device
// buffers of `a` / `b` on `copy_stream` while the persistent sweep kernel already runs on `stream` and
// waits, core by core, for the data (per-core ready flags set by the copy stream).  Falls back to
// copy-then-compute when the fused kernel does not apply.  The result is complete when `stream` is.
size_t inner_streamed_workspace_bytes(const TTDesc& a, const TTDesc& b);
int inner_streamed(const TTDesc& a, const TTDesc& b, const double* const* a_host, const double* const* b_host,
                   double* out_dev, void* ws, size_t ws_bytes, cudaStream_t stream, cudaStream_t copy_stream);

// --- dense contraction of a chain (inner.cu) ---
int tt_to_dense(const TTDesc& a, double* out_dev, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t tt_to_dense_workspace_bytes(const TTDesc& a);

}  // namespace ttb
