// Sequence pools: the inserted bases of a table's INS rows, copied out of the 4-bit query sequences into a
// compact buffer that travels with the table.
//
// Why: compute_distance needs candidate.sequence of both haplotypes (reference SVIM_COMBINE.py:70-75).  The full
// query sequences (0.65 GB per haplotype for a human assembly) only have to be on the device that ran the scan;
// what pairing needs is a few MB.  A pooled table can be all-gathered between ranks, uploaded from the host in a
// few hundred microseconds, and paired without the record image it came from.
#include <algorithm>

#include "pairing.cuh"

namespace {

__global__ void pool_sizes_kernel(const svb_row* __restrict__ rows, uint32_t n, uint32_t* __restrict__ bytes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bytes[i] = rows[i].type == SVB_INS ? (rows[i].seq_len + 1u) / 2u : 0u;
}

// one warp per row: re-pack seq_len nibbles starting at an arbitrary nibble offset to a byte-aligned run
__global__ void pool_gather_kernel(const svb_row* __restrict__ rows, uint32_t n, const uint32_t* __restrict__ off32,
                                   const uint8_t* __restrict__ seq4, const uint64_t* __restrict__ seq_off,
                                   uint8_t* __restrict__ pool, uint64_t* __restrict__ pool_off) {
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (row > n) return;
    if (row == n) {
        if (lane == 0) pool_off[n] = off32[n];
        return;
    }
    if (lane == 0) pool_off[row] = off32[row];
    const svb_row r = rows[row];
    if (r.type != SVB_INS || r.seq_len == 0u) return;
    const uint64_t src = seq_off[r.aln_idx] * 2ull + r.seq_pos;
    const uint32_t nbytes = (r.seq_len + 1u) / 2u;
    uint8_t* dst = pool + off32[row];
    // four output bytes per lane and round: five independent source loads in flight instead of two dependent ones per byte
    // (a 10,000-base insertion is 40 rounds of its warp instead of 157; the source may be pinned host memory read over PCIe)
    const uint64_t first = src >> 1;
    const bool odd = (src & 1ull) != 0ull;
    for (uint32_t b0 = 4u * lane; b0 < nbytes; b0 += 128u) {
        uint32_t s[5];
#pragma unroll
        for (uint32_t i = 0; i < 5u; ++i) {
            const uint32_t b = b0 + i;
            s[i] = (b < nbytes + (odd ? 1u : 0u) && (i < 4u || odd)) ? seq4[first + b] : 0u;
        }
#pragma unroll
        for (uint32_t i = 0; i < 4u; ++i) {
            const uint32_t b = b0 + i;
            if (b >= nbytes) break;
            uint32_t v = odd ? (((s[i] & 15u) << 4) | (s[i + 1u] >> 4)) : s[i];
            if (2u * b + 1u >= r.seq_len) v &= 0xF0u;             // the last nibble of an odd length is zero
            dst[b] = static_cast<uint8_t>(v);
        }
    }
}

}  // namespace

int table_drop_pool(svb_table* t) {
    if (t->d_pool) cudaFreeAsync(t->d_pool, t->stream);
    if (t->d_pool_off) cudaFreeAsync(t->d_pool_off, t->stream);
    t->d_pool = nullptr;
    t->d_pool_off = nullptr;
    t->pool_bytes = 0;
    return SVB_OK;
}

extern "C" {

// sizes -> exclusive scan -> one warp per row copies its nibbles; seq4 / seq_off may live in HBM or in pinned host memory
static int gather_pool(svb_ctx* ctx, svb_table* t, const uint8_t* seq4, const uint64_t* seq_off) {
    table_drop_pool(t);
    const uint32_t n = static_cast<uint32_t>(t->n);
    uint32_t* off32 = nullptr;
    SVB_CUDA(ctx, cudaMallocAsync(&off32, sizeof(uint32_t) * (static_cast<size_t>(n) + 2), ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&t->d_pool_off, sizeof(uint64_t) * (static_cast<size_t>(n) + 1), ctx->stream));
    if (n) {
        pool_sizes_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(t->d_rows, n, off32);
        ctx->launches += 1;
    }
    int rc = launch_scan_u32(ctx, off32, n, ctx->d_counters + 10);
    if (rc != SVB_OK) return rc;
    SVB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 10, ctx->d_counters + 10, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    t->pool_bytes = ctx->h_pinned[10];
    SVB_CUDA(ctx, cudaMallocAsync(&t->d_pool, std::max<uint64_t>(t->pool_bytes, 1), ctx->stream));
    const uint64_t threads = (static_cast<uint64_t>(n) + 1) * 32;
    pool_gather_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, ctx->stream>>>(t->d_rows, n, off32, seq4, seq_off,
                                                                                              t->d_pool, t->d_pool_off);
    ctx->launches += 1;
    SVB_CUDA(ctx, cudaGetLastError());
    SVB_CUDA(ctx, cudaFreeAsync(off32, ctx->stream));
    return SVB_OK;
}

}  // extern "C"

namespace {
__global__ void pool_check_kernel(const uint32_t* __restrict__ off32, uint32_t n, unsigned long long expected, uint32_t* dev_status) {
    if (off32[n] != expected) atomicOr(dev_status, DEV_ERR_CAPACITY);
}
}  // namespace

// The same gather when the pool's size is already known (svb_collect2 sums the inserted bytes on the device before its one
// synchronisation): nothing waits.  A total that disagrees with the rows raises the device status word.
int gather_pool_known(svb_ctx* ctx, svb_table* t, const uint8_t* seq4, const uint64_t* seq_off, uint64_t pool_bytes) {
    table_drop_pool(t);
    const uint32_t n = static_cast<uint32_t>(t->n);
    uint32_t* off32 = nullptr;
    SVB_CUDA(ctx, cudaMallocAsync(&off32, sizeof(uint32_t) * (static_cast<size_t>(n) + 2), ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&t->d_pool_off, sizeof(uint64_t) * (static_cast<size_t>(n) + 1), ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&t->d_pool, std::max<uint64_t>(pool_bytes, 1), ctx->stream));
    t->pool_bytes = pool_bytes;
    if (n) {
        pool_sizes_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(t->d_rows, n, off32);
        ctx->launches += 1;
    }
    int rc = launch_scan_u32(ctx, off32, n, ctx->d_counters + (ctx->stream == ctx->side ? 26 : 10));     // (two gathers may run side by side)
    if (rc != SVB_OK) return rc;
    pool_check_kernel<<<1, 1, 0, ctx->stream>>>(off32, n, pool_bytes, ctx->d_status);
    const uint64_t threads = (static_cast<uint64_t>(n) + 1) * 32;
    pool_gather_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, ctx->stream>>>(t->d_rows, n, off32, seq4, seq_off,
                                                                                              t->d_pool, t->d_pool_off);
    ctx->launches += 2;
    SVB_CUDA(ctx, cudaGetLastError());
    SVB_CUDA(ctx, cudaFreeAsync(off32, ctx->stream));
    return SVB_OK;
}

extern "C" {

// device address of a host pointer the GPU can read in place (pinned / registered memory), else nullptr
static const void* device_view_of_host(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if ((attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged) && attr.devicePointer) return attr.devicePointer;
    return nullptr;
}

int svb_table_gather_sequences(svb_ctx* ctx, svb_table* t, const svb_records* rec) {
    if (!ctx || !t || !rec) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_table_gather_sequences") : SVB_ERR_ARG;
    if (!rec->d_seq_off) return svb_fail(ctx, SVB_ERR_ARG, "svb_table_gather_sequences: call svb_records_set_sequences first");
    cudaSetDevice(ctx->device);
    return gather_pool(ctx, t, rec->d_seq4, rec->d_seq_off);
}

int svb_table_attach_sequences_host(svb_ctx* ctx, svb_table* t, const uint8_t* seq4, const uint64_t* seq_off) {
    if (!ctx || !t || !seq_off) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_table_attach_sequences_host") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    // pinned host buffers: the gather kernel reads the inserted bases in place over PCIe (a few MB of the 0.65 GB), no
    // host pass and no staging copy.  Pageable buffers take the host gather below.
    if (seq4) {
        const uint8_t* dv_seq = static_cast<const uint8_t*>(device_view_of_host(seq4));
        const uint64_t* dv_off = static_cast<const uint64_t*>(device_view_of_host(seq_off));
        if (dv_seq && dv_off) return gather_pool(ctx, t, dv_seq, dv_off);
    }
    table_drop_pool(t);
    const uint64_t n = t->n;
    std::vector<svb_row> rows(n);
    if (n) SVB_CUDA(ctx, cudaMemcpyAsync(rows.data(), t->d_rows, sizeof(svb_row) * n, cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<uint64_t> off(n + 1, 0);
    for (uint64_t i = 0; i < n; ++i) off[i + 1] = off[i] + (rows[i].type == SVB_INS ? (rows[i].seq_len + 1u) / 2u : 0u);
    std::vector<uint8_t> pool(off[n]);
    for (uint64_t i = 0; i < n; ++i) {
        const svb_row& r = rows[i];
        if (r.type != SVB_INS || !r.seq_len) continue;
        const uint64_t src = seq_off[r.aln_idx] * 2ull + r.seq_pos;
        uint8_t* dst = pool.data() + off[i];
        if ((src & 1ull) == 0) {
            memcpy(dst, seq4 + (src >> 1), (r.seq_len + 1u) / 2u);
            if (r.seq_len & 1u) dst[r.seq_len / 2u] &= 0xF0;
        } else {
            for (uint32_t b = 0; b < (r.seq_len + 1u) / 2u; ++b) {
                const uint64_t n0 = src + 2ull * b, n1 = n0 + 1ull;
                const uint32_t hi = seq4[n0 >> 1] & 15u;
                const uint32_t lo = (2u * b + 1u < r.seq_len) ? (seq4[n1 >> 1] >> 4) : 0u;
                dst[b] = static_cast<uint8_t>((hi << 4) | lo);
            }
        }
    }
    t->pool_bytes = off[n];
    SVB_CUDA(ctx, cudaMallocAsync(&t->d_pool, std::max<uint64_t>(t->pool_bytes, 1), ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&t->d_pool_off, sizeof(uint64_t) * (n + 1), ctx->stream));
    if (t->pool_bytes) SVB_CUDA(ctx, cudaMemcpyAsync(t->d_pool, pool.data(), t->pool_bytes, cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(t->d_pool_off, off.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SVB_OK;
}

int svb_table_pool_to_host(svb_ctx* ctx, const svb_table* t, uint8_t* pool_dst, uint64_t cap_bytes, uint64_t* off_dst,
                           uint64_t* pool_bytes) {
    if (!ctx || !t || !pool_bytes) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_table_pool_to_host") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    *pool_bytes = t->pool_bytes;
    if (!t->d_pool_off) return svb_fail(ctx, SVB_ERR_ARG, "svb_table_pool_to_host: the table has no sequence pool");
    if (off_dst) SVB_CUDA(ctx, cudaMemcpyAsync(off_dst, t->d_pool_off, sizeof(uint64_t) * (t->n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (pool_dst && t->pool_bytes) {
        if (cap_bytes < t->pool_bytes) return svb_fail(ctx, SVB_ERR_ARG, "svb_table_pool_to_host: destination too small");
        SVB_CUDA(ctx, cudaMemcpyAsync(pool_dst, t->d_pool, t->pool_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SVB_OK;
}

int svb_table_set_pool_from_host(svb_ctx* ctx, svb_table* t, const uint8_t* pool, uint64_t pool_bytes, const uint64_t* row_off) {
    if (!ctx || !t || (t->n && !row_off) || (pool_bytes && !pool)) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_table_set_pool_from_host") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    table_drop_pool(t);
    t->pool_bytes = pool_bytes;
    SVB_CUDA(ctx, cudaMallocAsync(&t->d_pool, std::max<uint64_t>(t->pool_bytes, 1), ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&t->d_pool_off, sizeof(uint64_t) * (t->n + 1), ctx->stream));
    if (t->pool_bytes) SVB_CUDA(ctx, cudaMemcpyAsync(t->d_pool, pool, t->pool_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (t->n) SVB_CUDA(ctx, cudaMemcpyAsync(t->d_pool_off, row_off, sizeof(uint64_t) * t->n, cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SVB_OK;
}

}  // extern "C"
