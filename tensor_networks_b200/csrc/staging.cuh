// Host -> device staging of pageable memory through a pinned ring; see staging.cu.
#pragma once

#include <vector>

#include "common.cuh"

namespace ttb {

struct HostCopy {
    void* dst;        // device
    const void* src;  // host, pageable or pinned
    size_t bytes;
    int flag_after;   // >= 0: flags_dev[flag_after] is set (stream-ordered) after this copy; -1: none
};

// Enqueue the copies on `stream` in the given order.  Pinned sources are enqueued directly; pageable
// ones are staged chunk by chunk by a pool of host threads.  Returns when every chunk has been ENQUEUED
// (the copies themselves may still be in flight).
int staged_h2d(const std::vector<HostCopy>& copies, int* flags_dev, cudaStream_t stream);

}  // namespace ttb
