// Host -> device staging for callers that hold PAGEABLE memory (numpy arrays): the drop-in path.
//
// A pytens user calls TensorNetwork.inner() on cores that live in ordinary numpy buffers
// (pytens/algs.py:585-587).  cudaMemcpyAsync from pageable memory is staged by the driver through
// one internal buffer by one thread and reaches a fraction of the PCIe / C2C rate.  Here a small pool
// of host threads copies 4 MB chunks into a ring of pinned slots (several memcpy streams in
// parallel saturate the host memory system) and each chunk is handed to the copy engine with
// cudaMemcpyAsync as soon as it is staged -- in the original order, so that the per-core "ready"
// flags of the streamed sweep kernel keep their meaning.  Sources that already are pinned skip the
// staging and are enqueued directly.
#include "staging.cuh"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace ttb {

namespace {

// Ring geometry: TTB_STAGE_SLOT_MB (default 4) x TTB_STAGE_SLOTS (default 24, at most kMaxSlots).
constexpr int kMaxSlots = 64;
size_t env_slot_bytes() {
    const char* e = getenv("TTB_STAGE_SLOT_MB");
    const int mb = e ? atoi(e) : 0;
    return size_t(mb > 0 && mb <= 256 ? mb : 4) << 20;
}
int env_slots() {
    const char* e = getenv("TTB_STAGE_SLOTS");
    const int n = e ? atoi(e) : 0;
    return (n >= 2 && n <= kMaxSlots) ? n : 24;
}
const size_t kSlotBytes = env_slot_bytes();
const int kSlots = env_slots();

// Copy into the pinned ring.  TTB_STAGE_NT=1 uses streaming (non-temporal) stores: the destination is read
// next by the copy engine, never by this core, so it need not be allocated in (or read into) the cache.
const bool kNtStores = [] {
    const char* e = getenv("TTB_STAGE_NT");
    return e != nullptr && e[0] == '1';
}();
void stage_copy(char* dst, const char* src, size_t bytes) {
#if defined(__x86_64__)
    if (kNtStores && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        const size_t n16 = bytes / 64;
        const __m128i* s = reinterpret_cast<const __m128i*>(src);
        __m128i* d = reinterpret_cast<__m128i*>(dst);
        for (size_t i = 0; i < n16; ++i) {
            const __m128i a = _mm_load_si128(s + 4 * i), b = _mm_load_si128(s + 4 * i + 1);
            const __m128i c = _mm_load_si128(s + 4 * i + 2), e = _mm_load_si128(s + 4 * i + 3);
            _mm_stream_si128(d + 4 * i, a);
            _mm_stream_si128(d + 4 * i + 1, b);
            _mm_stream_si128(d + 4 * i + 2, c);
            _mm_stream_si128(d + 4 * i + 3, e);
        }
        _mm_sfence();
        if (bytes % 64) std::memcpy(dst + n16 * 64, src + n16 * 64, bytes % 64);
        return;
    }
#endif
    std::memcpy(dst, src, bytes);
}

struct Chunk {
    void* dst;
    const char* src;
    size_t bytes;
    int flag_after;  // >= 0: set flags[flag_after] once this chunk (and everything before it) has landed
    bool pinned;
};

struct Job {
    const std::vector<Chunk>* chunks = nullptr;
    int* flags_dev = nullptr;
    cudaStream_t stream = nullptr;
    std::atomic<size_t> next_take{0};   // next chunk to stage
    std::atomic<size_t> next_issue{0};  // next chunk allowed to be enqueued (keeps stream order)
    std::atomic<int> error{0};
    std::atomic<int> active{0};
};

class Stager {
  public:
    static Stager& get() {
        static Stager s;
        return s;
    }

    int run(const std::vector<Chunk>& chunks, int* flags_dev, cudaStream_t stream) {
        std::lock_guard<std::mutex> call_guard(call_mu_);  // one staged transfer at a time per process
        if (!ensure_ring()) return kCudaError;
        Job job;
        job.chunks = &chunks;
        job.flags_dev = flags_dev;
        job.stream = stream;
        job.active = nthreads_;
        cudaGetDevice(&device_);
        {
            std::lock_guard<std::mutex> g(mu_);
            job_ = &job;
            ++generation_;
        }
        cv_.notify_all();
        work(job);  // the calling thread stages as well
        {
            std::unique_lock<std::mutex> g(mu_);
            done_cv_.wait(g, [&] { return job.active.load() == 0; });
            job_ = nullptr;
        }
        if (job.error.load()) {
            set_last_error("staged host->device copy failed (cudaMemcpyAsync / event)");
            return kCudaError;
        }
        return kOk;
    }

  private:
    Stager() {
        const char* e = getenv("TTB_STAGE_THREADS");
        int want = e ? atoi(e) : 0;
        if (want <= 0) want = int(std::min<unsigned>(8u, std::max(2u, std::thread::hardware_concurrency() / 2)));
        want = std::min(want, 64);
        nthreads_ = std::max(1, want);
        for (int i = 0; i < nthreads_; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~Stager() {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
        // the pinned ring is left to process teardown (the CUDA context may already be gone)
    }

    bool ensure_ring() {
        if (ring_) return true;
        void* p = nullptr;
        if (cudaHostAlloc(&p, kSlotBytes * kSlots, cudaHostAllocDefault) != cudaSuccess) {
            set_last_error("staging: cudaHostAlloc of the pinned ring failed");
            return false;
        }
        ring_ = static_cast<char*>(p);
        for (int i = 0; i < kSlots; ++i) {
            if (cudaEventCreateWithFlags(&slot_free_[i], cudaEventDisableTiming) != cudaSuccess) return false;
            slot_used_[i] = false;
        }
        return true;
    }

    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            Job* job = nullptr;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return stop_ || (job_ != nullptr && generation_ != seen); });
                if (stop_) return;
                seen = generation_;
                job = job_;
            }
            cudaSetDevice(device_);
            work(*job);
            if (job->active.fetch_sub(1) == 1) {
                std::lock_guard<std::mutex> g(mu_);
                done_cv_.notify_all();
            }
        }
    }

    // Take chunks in order; stage into the slot (chunk index mod kSlots) once the copy that last used the
    // slot has finished; enqueue strictly in chunk order.
    void work(Job& job) {
        const std::vector<Chunk>& cs = *job.chunks;
        for (;;) {
            const size_t i = job.next_take.fetch_add(1);
            if (i >= cs.size()) return;
            const Chunk& c = cs[i];
            const int slot = int(i % kSlots);
            const void* from = c.src;
            if (!c.pinned) {
                // slot reuse: chunk i - kSlots was enqueued before (next_issue is monotone), wait for its copy
                if (i >= size_t(kSlots)) {
                    while (job.next_issue.load(std::memory_order_acquire) <= i - kSlots && !job.error.load()) std::this_thread::yield();
                    if (cudaEventSynchronize(slot_free_[slot]) != cudaSuccess) job.error = 1;
                } else if (slot_used_[slot]) {
                    if (cudaEventSynchronize(slot_free_[slot]) != cudaSuccess) job.error = 1;
                }
                char* to = ring_ + size_t(slot) * kSlotBytes;
                stage_copy(to, c.src, c.bytes);
                from = to;
            }
            while (job.next_issue.load(std::memory_order_acquire) != i && !job.error.load()) std::this_thread::yield();
            if (!job.error.load()) {
                if (c.bytes && cudaMemcpyAsync(c.dst, from, c.bytes, cudaMemcpyHostToDevice, job.stream) != cudaSuccess) job.error = 1;
                if (!c.pinned) {
                    if (cudaEventRecord(slot_free_[slot], job.stream) != cudaSuccess) job.error = 1;
                    slot_used_[slot] = true;
                }
                if (c.flag_after >= 0 && job.flags_dev != nullptr)
                    if (cudaMemsetAsync(job.flags_dev + c.flag_after, 1, sizeof(int), job.stream) != cudaSuccess) job.error = 1;
            }
            job.next_issue.store(i + 1, std::memory_order_release);
        }
    }

    std::mutex call_mu_, mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<std::thread> workers_;
    Job* job_ = nullptr;
    unsigned long long generation_ = 0;
    bool stop_ = false;
    int nthreads_ = 1;
    int device_ = 0;
    char* ring_ = nullptr;
    cudaEvent_t slot_free_[kMaxSlots];
    bool slot_used_[kMaxSlots];
};

bool is_pinned(const void* p) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

}  // namespace

int staged_h2d(const std::vector<HostCopy>& copies, int* flags_dev, cudaStream_t stream) {
    std::vector<Chunk> chunks;
    bool any_pageable = false;
    for (const HostCopy& c : copies) {
        const bool pinned = is_pinned(c.src);
        any_pageable = any_pageable || !pinned;
        const char* src = static_cast<const char*>(c.src);
        char* dst = static_cast<char*>(c.dst);
        size_t left = c.bytes;
        if (left == 0 && c.flag_after >= 0) chunks.push_back({dst, src, 0, c.flag_after, true});
        while (left > 0) {
            const size_t n = pinned ? left : std::min(left, kSlotBytes);
            chunks.push_back({dst, src, n, (n == left) ? c.flag_after : -1, pinned});
            src += n;
            dst += n;
            left -= n;
        }
    }
    if (!any_pageable) {  // nothing to stage: plain enqueue
        for (const Chunk& c : chunks) {
            if (c.bytes) TTB_CHECK_CUDA(cudaMemcpyAsync(c.dst, c.src, c.bytes, cudaMemcpyHostToDevice, stream));
            if (c.flag_after >= 0 && flags_dev) TTB_CHECK_CUDA(cudaMemsetAsync(flags_dev + c.flag_after, 1, sizeof(int), stream));
        }
        return kOk;
    }
    return Stager::get().run(chunks, flags_dev, stream);
}

}  // namespace ttb
