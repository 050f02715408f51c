// Multi-GPU exchange of candidate tables, device resident.
//
// The reference has one process and one list per candidate type (SVIM_COLLECT.py:67-91), so "every candidate is
// visible to pair_candidates" is free there.  With the records sharded by contig over ranks, walk-derived
// candidates can land on a contig another rank owns and both haplotypes of a key contig must meet on one rank.
// svb_exchange_pack lays both haplotype tables (rows, pool offsets, sequence pool) of this rank into ONE device
// buffer that the caller all-gathers (NCCL, no host staging); svb_exchange_unpack turns the gathered buffers into
// the table this rank pairs: rows of all ranks in the global append order (ordinal), restricted to the key contigs
// this rank owns, with one concatenated pool.  Every rank's rows already are in ordinal order, so the global order is
// a W-way merge: a row's position is the sum over the runs of "rows before me" (binary search), counted over owned
// rows only through an exclusive scan of the ownership flags.
#include <algorithm>
#include <vector>

#include "pairing.cuh"

int table_drop_pool(svb_table* t);

namespace {

constexpr int EXCH_MAX_WORLD = 16;

struct Layout {                      // byte offsets of one rank's sections inside its packed buffer
    uint64_t rows[2], off[2], pool[2], total;
};

inline uint64_t align16(uint64_t v) { return (v + 15ull) & ~15ull; }

Layout layout_of(const uint64_t sizes[4]) {
    Layout l;
    uint64_t at = 0;
    for (int h = 0; h < 2; ++h) {
        const uint64_t n = sizes[2 * h], pool = sizes[2 * h + 1];
        l.rows[h] = at;
        at += n * sizeof(svb_row);
        l.off[h] = at;
        at = align16(at + (n + 1) * sizeof(uint64_t));
        l.pool[h] = at;
        at = align16(at + pool);
    }
    l.total = at;
    return l;
}

struct Run {
    const svb_row* rows;
    const uint64_t* off;             // pool offsets relative to this run's pool (or nullptr)
    uint32_t n, row_base;
    uint64_t pool_base;
};

struct Runs {
    Run run[EXCH_MAX_WORLD];
    int world;
    uint32_t total;
};

__device__ __forceinline__ int32_t key_contig(const svb_row& r) {
    // Candidate.get_key (reference SVCandidate.py:17-19,147-148,292-293,386-387): INS and DUP_INT key on the destination
    return (r.type == SVB_INS || r.type == SVB_DUP_INT) ? r.dst_tid : r.src_tid;
}

__device__ __forceinline__ void locate(const Runs& rs, uint32_t x, int& r, uint32_t& i) {
    r = 0;
    while (r + 1 < rs.world && x >= rs.run[r + 1].row_base) ++r;
    i = x - rs.run[r].row_base;
}

__global__ void owned_flags_kernel(const __grid_constant__ Runs rs, const int32_t* __restrict__ owner, int n_contig, int rank,
                                   uint32_t* __restrict__ flags) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x > rs.total) return;
    if (x == rs.total) { flags[x] = 0u; return; }
    int r;
    uint32_t i;
    locate(rs, x, r, i);
    const int32_t tid = key_contig(rs.run[r].rows[i]);
    flags[x] = (tid >= 0 && tid < n_contig && owner[tid] == rank) ? 1u : 0u;
}

__global__ void place_rows_kernel(const __grid_constant__ Runs rs, const uint32_t* __restrict__ scanned, svb_row* __restrict__ out_rows,
                                  uint64_t* __restrict__ out_off) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= rs.total || scanned[x + 1] == scanned[x]) return;        // not owned
    int r;
    uint32_t i;
    locate(rs, x, r, i);
    const svb_row row = rs.run[r].rows[i];
    uint32_t pos = 0;
    for (int q = 0; q < rs.world; ++q) {
        const Run& run = rs.run[q];
        uint32_t lo = 0, hi = run.n;                                   // rows of run q with a smaller ordinal
        if (q == r) lo = hi = i;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (run.rows[mid].ordinal < row.ordinal) lo = mid + 1; else hi = mid;
        }
        pos += scanned[run.row_base + lo] - scanned[run.row_base];
    }
    out_rows[pos] = row;
    if (out_off) out_off[pos] = rs.run[r].off ? rs.run[r].pool_base + rs.run[r].off[i] : 0ull;
}

__global__ void remap_records_kernel(svb_row* __restrict__ rows, uint32_t n, const uint32_t* __restrict__ global_idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t g = global_idx[rows[i].aln_idx];
    rows[i].aln_idx = static_cast<uint32_t>(g);
    rows[i].ordinal = (g << 32) | (rows[i].ordinal & 0xFFFFFFFFull);
}

}  // namespace

extern "C" {

void* svb_stream(svb_ctx* ctx) { return ctx ? reinterpret_cast<void*>(ctx->stream) : nullptr; }

int svb_device_alloc(svb_ctx* ctx, uint64_t bytes, void** out) {
    if (!ctx || !out) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_device_alloc") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    SVB_CUDA(ctx, cudaMallocAsync(out, std::max<uint64_t>(bytes, 16), ctx->stream));
    return SVB_OK;
}

void svb_device_free(svb_ctx* ctx, void* p) {
    if (ctx && p) cudaFreeAsync(p, ctx->stream);
}

int svb_records_set_global_index(svb_ctx* ctx, svb_records* rec, const uint32_t* global_idx) {
    if (!ctx || !rec || (rec->n_aln && !global_idx)) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_records_set_global_index") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (rec->d_global_idx) cudaFreeAsync(rec->d_global_idx, rec->stream);
    SVB_CUDA(ctx, cudaMallocAsync(&rec->d_global_idx, sizeof(uint32_t) * std::max<uint64_t>(rec->n_aln, 1), ctx->stream));
    if (rec->n_aln) SVB_CUDA(ctx, cudaMemcpyAsync(rec->d_global_idx, global_idx, sizeof(uint32_t) * rec->n_aln, cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SVB_OK;
}

int svb_table_remap_records(svb_ctx* ctx, svb_table* t, const svb_records* rec) {
    if (!ctx || !t || !rec || !rec->d_global_idx) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_table_remap_records: no global index") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (!t->n) return SVB_OK;
    remap_records_kernel<<<static_cast<unsigned>((t->n + 255) / 256), 256, 0, ctx->stream>>>(t->d_rows, static_cast<uint32_t>(t->n), rec->d_global_idx);
    ctx->launches += 1;
    SVB_CUDA(ctx, cudaGetLastError());
    return SVB_OK;
}

uint64_t svb_exchange_bytes(const uint64_t sizes[4]) { return sizes ? layout_of(sizes).total : 0; }

int svb_exchange_sizes(const svb_table* t1, const svb_table* t2, uint64_t sizes[4]) {
    if (!t1 || !t2 || !sizes) return SVB_ERR_ARG;
    sizes[0] = t1->n; sizes[1] = t1->d_pool_off ? t1->pool_bytes : 0;
    sizes[2] = t2->n; sizes[3] = t2->d_pool_off ? t2->pool_bytes : 0;
    return SVB_OK;
}

int svb_exchange_pack(svb_ctx* ctx, const svb_table* t1, const svb_table* t2, void* d_buf, uint64_t cap_bytes) {
    if (!ctx || !t1 || !t2 || !d_buf) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_pack") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    uint64_t sizes[4];
    svb_exchange_sizes(t1, t2, sizes);
    const Layout l = layout_of(sizes);
    if (l.total > cap_bytes) return svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_pack: buffer too small");
    uint8_t* base = static_cast<uint8_t*>(d_buf);
    const svb_table* t[2] = {t1, t2};
    for (int h = 0; h < 2; ++h) {
        if (t[h]->n) SVB_CUDA(ctx, cudaMemcpyAsync(base + l.rows[h], t[h]->d_rows, sizeof(svb_row) * t[h]->n, cudaMemcpyDeviceToDevice, ctx->stream));
        if (t[h]->d_pool_off) {
            SVB_CUDA(ctx, cudaMemcpyAsync(base + l.off[h], t[h]->d_pool_off, sizeof(uint64_t) * t[h]->n, cudaMemcpyDeviceToDevice, ctx->stream));
            if (t[h]->pool_bytes) SVB_CUDA(ctx, cudaMemcpyAsync(base + l.pool[h], t[h]->d_pool, t[h]->pool_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        } else {
            SVB_CUDA(ctx, cudaMemsetAsync(base + l.off[h], 0, sizeof(uint64_t) * (t[h]->n + 1), ctx->stream));
        }
    }
    return SVB_OK;
}

int svb_exchange_unpack(svb_ctx* ctx, const void* d_gathered, uint64_t stride, const uint64_t* sizes, int world, int hap,
                        const int32_t* owner, int n_contig, int rank, svb_table** out) {
    if (!ctx || !d_gathered || !sizes || !owner || !out || world < 1 || world > EXCH_MAX_WORLD || hap < 1 || hap > 2)
        return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_unpack") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    const int h = hap - 1;
    const uint8_t* base = static_cast<const uint8_t*>(d_gathered);
    Runs rs;
    memset(&rs, 0, sizeof rs);
    rs.world = world;
    uint64_t rows_total = 0, pool_total = 0;
    std::vector<Layout> lay(world);
    for (int r = 0; r < world; ++r) {
        lay[r] = layout_of(sizes + 4 * r);
        if (lay[r].total > stride) return svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_unpack: stride smaller than a rank's payload");
        const uint8_t* p = base + stride * r;
        rs.run[r].rows = reinterpret_cast<const svb_row*>(p + lay[r].rows[h]);
        rs.run[r].off = reinterpret_cast<const uint64_t*>(p + lay[r].off[h]);
        rs.run[r].n = static_cast<uint32_t>(sizes[4 * r + 2 * h]);
        rs.run[r].row_base = static_cast<uint32_t>(rows_total);
        rs.run[r].pool_base = pool_total;
        rows_total += sizes[4 * r + 2 * h];
        pool_total += sizes[4 * r + 2 * h + 1];
    }
    if (rows_total >= 0xFFFFFFF0ull) return svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_unpack: too many rows");
    rs.total = static_cast<uint32_t>(rows_total);
    svb_table* t = new (std::nothrow) svb_table();
    if (!t) return svb_fail(ctx, SVB_ERR_NOMEM, "svb_exchange_unpack");
    t->stream = ctx->stream;
    t->cap = std::max<uint64_t>(rows_total, 1);
    int32_t* d_owner = nullptr;
    uint32_t* d_flags = nullptr;
    auto fail = [&](int rc) { svb_table_free(t); return rc; };
#define EX_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(svb_fail(ctx, SVB_ERR_CUDA, cudaGetErrorString(e_))); } while (0)
    EX_CUDA(cudaMallocAsync(&t->d_rows, sizeof(svb_row) * t->cap, ctx->stream));
    EX_CUDA(cudaMallocAsync(&t->d_pool_off, sizeof(uint64_t) * (t->cap + 1), ctx->stream));
    EX_CUDA(cudaMallocAsync(&t->d_pool, std::max<uint64_t>(pool_total, 1), ctx->stream));
    EX_CUDA(cudaMallocAsync(&d_owner, sizeof(int32_t) * std::max(n_contig, 1), ctx->stream));
    EX_CUDA(cudaMallocAsync(&d_flags, sizeof(uint32_t) * (rows_total + 2), ctx->stream));
    t->pool_bytes = pool_total;
    if (n_contig) EX_CUDA(cudaMemcpyAsync(d_owner, owner, sizeof(int32_t) * n_contig, cudaMemcpyHostToDevice, ctx->stream));
    for (int r = 0; r < world; ++r) {
        const uint64_t bytes = sizes[4 * r + 2 * h + 1];
        if (bytes) EX_CUDA(cudaMemcpyAsync(t->d_pool + rs.run[r].pool_base, base + stride * r + lay[r].pool[h], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    const unsigned blocks = static_cast<unsigned>((rows_total + 1 + 255) / 256);
    owned_flags_kernel<<<blocks, 256, 0, ctx->stream>>>(rs, d_owner, n_contig, rank, d_flags);
    ctx->launches += 1;
    int rc = launch_scan_u32(ctx, d_flags, static_cast<uint32_t>(rows_total + 1), ctx->d_counters + 11);
    if (rc != SVB_OK) return fail(rc);
    if (rows_total) {
        place_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(rs, d_flags, t->d_rows, t->d_pool_off);
        ctx->launches += 1;
    }
    EX_CUDA(cudaGetLastError());
    EX_CUDA(cudaMemcpyAsync(ctx->h_pinned + 11, ctx->d_counters + 11, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    EX_CUDA(cudaStreamSynchronize(ctx->stream));
#undef EX_CUDA
    t->n = ctx->h_pinned[11];
    cudaFreeAsync(d_owner, ctx->stream);
    cudaFreeAsync(d_flags, ctx->stream);
    *out = t;
    return SVB_OK;
}

const void* svb_table_device_rows(const svb_table* t) { return t ? t->d_rows : nullptr; }

}  // extern "C"
