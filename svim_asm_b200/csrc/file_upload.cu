// File bytes -> device memory at PCIe speed.
//
// cudaMemcpy from a mapping of the page cache is a pageable copy: the driver stages it through one thread, about
// 11 GB/s measured here -- 0.29 s for a 3.1 GB FASTA, a third of a whole file -> variants.vcf run.  Here several host
// threads pread() 4 MB chunks of the file straight into their own pinned staging slots and send each with an
// asynchronous copy on their own stream, so the page-cache reads run in parallel and overlap the transfers.
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

// Uploads in flight that something is waiting for right now (a BAM file: its inflate kernel cannot start before the bytes are
// there).  A background upload (the FASTA, needed only when pairing starts) pauses between chunks while this is non-zero: the
// two would otherwise share the page-cache reads and the link half and half, and the ingest would wait twice as long.
std::atomic<int> g_urgent_uploads(0);

constexpr int UP_MAX_THREADS = 8;
constexpr size_t UP_SLOT = 4u << 20;

struct Uploader {
    int n_threads = 0;
    uint8_t* pinned = nullptr;                       // n_threads x 2 slots
    cudaStream_t stream[UP_MAX_THREADS] = {};
    cudaEvent_t sent[UP_MAX_THREADS][2] = {};
};

Uploader* uploader_of(svb_ctx* ctx) {
    if (ctx->uploader) return static_cast<Uploader*>(ctx->uploader);
    Uploader* u = new (std::nothrow) Uploader();
    if (!u) return nullptr;
    const unsigned hw = std::thread::hardware_concurrency();
    u->n_threads = static_cast<int>(std::min<unsigned>(UP_MAX_THREADS, std::max(2u, hw / 2u)));
    bool ok = cudaHostAlloc(&u->pinned, UP_SLOT * 2 * static_cast<size_t>(u->n_threads), cudaHostAllocDefault) == cudaSuccess;
    for (int t = 0; ok && t < u->n_threads; ++t) {
        ok = cudaStreamCreateWithFlags(&u->stream[t], cudaStreamNonBlocking) == cudaSuccess;
        for (int s = 0; ok && s < 2; ++s) ok = cudaEventCreateWithFlags(&u->sent[t][s], cudaEventDisableTiming) == cudaSuccess;
    }
    ctx->uploader = u;
    if (!ok) {                                       // half built: give everything back, the next call tries again
        cudaGetLastError();
        upload_release(ctx);
        return nullptr;
    }
    return u;
}

}  // namespace

void upload_release(svb_ctx* ctx) {
    Uploader* u = static_cast<Uploader*>(ctx->uploader);
    if (!u) return;
    for (int t = 0; t < UP_MAX_THREADS; ++t) {
        for (int s = 0; s < 2; ++s)
            if (u->sent[t][s]) cudaEventDestroy(u->sent[t][s]);
        if (u->stream[t]) cudaStreamDestroy(u->stream[t]);
    }
    if (u->pinned) cudaFreeHost(u->pinned);
    delete u;
    ctx->uploader = nullptr;
}

// bytes [offset, offset + n) of the open file -> d_dst.  On return the bytes are in device memory (every worker has
// synchronised its stream), so work enqueued on ctx->stream afterwards sees them.
int upload_file_range(svb_ctx* ctx, int fd, uint64_t offset, uint64_t n, void* d_dst, bool wait_for_stream, bool background) {
    if (n == 0) return SVB_OK;
    Uploader* u = uploader_of(ctx);
    if (!u) return wait_for_stream ? svb_fail(ctx, SVB_ERR_NOMEM, "upload_file_range: staging buffers") : SVB_ERR_NOMEM;
    if (wait_for_stream) SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // d_dst may come from a stream-ordered allocation on ctx->stream
    const uint64_t n_chunks = (n + UP_SLOT - 1) / UP_SLOT;
    std::atomic<uint64_t> next(0);
    std::atomic<int> failed(0);                                  // 1 read error, 2 CUDA error
    const int device = ctx->device;
    struct Urgent {
        bool on;
        explicit Urgent(bool o) : on(o) { if (on) g_urgent_uploads.fetch_add(1); }
        ~Urgent() { if (on) g_urgent_uploads.fetch_sub(1); }
    } urgent(!background);
    auto work = [&](int t) {
        cudaSetDevice(device);
        for (int turn = 0; !failed.load(std::memory_order_relaxed); ++turn) {
            if (background)
                for (int waited = 0; g_urgent_uploads.load(std::memory_order_relaxed) > 0 && waited < 20000; ++waited)      // (at most 4 s: never a deadlock)
                    std::this_thread::sleep_for(std::chrono::microseconds(200));
            const uint64_t c = next.fetch_add(1);
            if (c >= n_chunks) break;
            const int s = turn & 1;
            uint8_t* slot = u->pinned + UP_SLOT * (2 * static_cast<size_t>(t) + s);
            if (turn >= 2 && cudaEventSynchronize(u->sent[t][s]) != cudaSuccess) {
                failed.store(2);
                break;
            }
            const uint64_t lo = c * UP_SLOT, bytes = std::min<uint64_t>(UP_SLOT, n - lo);
            uint64_t got = 0;
            while (got < bytes) {
                const ssize_t r = pread(fd, slot + got, bytes - got, static_cast<off_t>(offset + lo + got));
                if (r <= 0) break;
                got += static_cast<uint64_t>(r);
            }
            if (got != bytes) {
                failed.store(1);
                break;
            }
            if (cudaMemcpyAsync(static_cast<uint8_t*>(d_dst) + lo, slot, bytes, cudaMemcpyHostToDevice, u->stream[t]) != cudaSuccess ||
                cudaEventRecord(u->sent[t][s], u->stream[t]) != cudaSuccess) {
                failed.store(2);
                break;
            }
        }
        if (cudaStreamSynchronize(u->stream[t]) != cudaSuccess) failed.store(2);
    };
    const int n_workers = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(u->n_threads), n_chunks));
    std::vector<std::thread> pool;
    for (int t = 1; t < n_workers; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    if (!wait_for_stream && failed.load()) return failed.load() == 1 ? SVB_ERR_IO : SVB_ERR_CUDA;
    if (failed.load() == 1) return svb_fail(ctx, SVB_ERR_IO, "upload_file_range: short read");
    if (failed.load() == 2) return svb_fail(ctx, SVB_ERR_CUDA, "upload_file_range", cudaGetLastError());
    return SVB_OK;
}
