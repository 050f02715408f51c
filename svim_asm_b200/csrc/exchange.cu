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
#include <cstdlib>
#include <cstring>
#include <new>
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

// enqueue only: the number of rows kept lands in h_pinned[13 + hap - 1] once the stream has run (the caller synchronises once
// for both haplotypes and then sets the tables' sizes)
static int exchange_unpack_enqueue(svb_ctx* ctx, const void* d_gathered, uint64_t stride, const uint64_t* sizes, int world, int hap,
                                   const int32_t* owner, int n_contig, int rank, svb_table** out) {
    if (!ctx || !d_gathered || !sizes || !owner || !out || world < 1 || world > EXCH_MAX_WORLD || hap < 1 || hap > 2)
        return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_unpack") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    const int slot = 13 + hap - 1;               // device counter / pinned word of this haplotype ([12] is the status word)
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
    int rc = launch_scan_u32(ctx, d_flags, static_cast<uint32_t>(rows_total + 1), ctx->d_counters + slot);
    if (rc != SVB_OK) return fail(rc);
    if (rows_total) {
        place_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(rs, d_flags, t->d_rows, t->d_pool_off);
        ctx->launches += 1;
    }
    EX_CUDA(cudaGetLastError());
    EX_CUDA(cudaMemcpyAsync(ctx->h_pinned + slot, ctx->d_counters + slot, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
#undef EX_CUDA
    cudaFreeAsync(d_owner, ctx->stream);
    cudaFreeAsync(d_flags, ctx->stream);
    *out = t;
    return SVB_OK;
}

int svb_exchange_unpack(svb_ctx* ctx, const void* d_gathered, uint64_t stride, const uint64_t* sizes, int world, int hap,
                        const int32_t* owner, int n_contig, int rank, svb_table** out) {
    int rc = exchange_unpack_enqueue(ctx, d_gathered, stride, sizes, world, hap, owner, n_contig, rank, out);
    if (rc != SVB_OK) return rc;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        svb_table_free(*out);
        *out = nullptr;
        return svb_fail(ctx, SVB_ERR_CUDA, "svb_exchange_unpack");
    }
    (*out)->n = ctx->h_pinned[13 + hap - 1];
    return SVB_OK;
}

const void* svb_table_device_rows(const svb_table* t) { return t ? t->d_rows : nullptr; }

}  // extern "C"

// =====================================================================================================================
// Exchange window over peer memory (SURVEY.md 8e, C1 "fused after K5"): the all-gatherv of the candidate tables and of the
// paired rows WITHOUT a collective library and without host-visible sizes in between.
//
// Every rank owns a window in its HBM (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by every peer):
//     in slot [world]   one packed pair of haplotype tables per rank: 64-byte header {sizes[4], payload bytes} + svb_exchange_pack layout
//     result slot [world]  one run of paired rows per rank: 64-byte header {rows} + rows
//     flags             one 32-bit epoch per (kind, writer rank), 128 bytes apart
// `put` is ONE kernel: it stores this rank's payload straight into its slot of EVERY peer's window over NVLink (16-byte
// stores, grid-stride), and the last block to finish raises this rank's flag in every window (system-scope fences on both
// sides).  `wait` is a one-warp kernel on the consumer's stream that spins on the flags of its own window; what follows on
// that stream reads the gathered data in place.  Epochs only grow, so nothing is reset between steps; a rank overwrites a
// peer's slot for step s + 1 only after it has seen that peer's result flag of step s, which the peer raises after it
// has read everything of step s.
namespace {

constexpr uint32_t XW_HEADER = 64;
constexpr uint32_t XW_FLAG_STRIDE = 128;
constexpr int XW_MAX_PIECES = 8;

struct PutPiece {
    const uint8_t* src;
    uint64_t bytes;
    uint64_t dst_off;                // offset inside the destination slot
};
struct PutArgs {
    uint8_t* peer_slot[EXCH_MAX_WORLD];          // this rank's slot in every rank's window (own window included)
    uint32_t* peer_flag[EXCH_MAX_WORLD];         // this rank's flag in every rank's window
    PutPiece piece[XW_MAX_PIECES];
    uint64_t header[8];                          // written to offset 0 of every slot
    int world, n_pieces;
    uint32_t epoch;
    unsigned int* done;                          // device counter of finished blocks (zeroed by the caller)
};

__global__ void __launch_bounds__(256) exchange_put_kernel(const __grid_constant__ PutArgs a) {
    const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x, stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (int p = 0; p < a.world; ++p) {
        uint8_t* slot = a.peer_slot[p];
        if (tid < 8) reinterpret_cast<uint64_t*>(slot)[tid] = a.header[tid];
        for (int k = 0; k < a.n_pieces; ++k) {
            const PutPiece pc = a.piece[k];
            const uint4* s4 = reinterpret_cast<const uint4*>(pc.src);
            uint4* d4 = reinterpret_cast<uint4*>(slot + pc.dst_off);
            const uint64_t n16 = pc.bytes / 16;
            for (uint64_t i = tid; i < n16; i += stride) d4[i] = s4[i];
            for (uint64_t i = n16 * 16 + tid; i < pc.bytes; i += stride) slot[pc.dst_off + i] = pc.src[i];
        }
    }
    // every store of this block is visible system-wide before the block counts itself done; the last block raises the flags
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(a.done, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (last && threadIdx.x < static_cast<unsigned>(a.world)) {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(a.peer_flag[threadIdx.x]) = a.epoch;
        __threadfence_system();
    }
}

// spin until every writer's flag in OUR window has reached `epoch` (or give up after `timeout_ns`: a peer died)
__global__ void exchange_wait_kernel(const uint32_t* flags, int world, uint32_t epoch, unsigned long long timeout_ns, uint32_t* dev_status) {
    const int p = threadIdx.x;
    if (p < world) {
        const volatile uint32_t* f = reinterpret_cast<const volatile uint32_t*>(reinterpret_cast<const uint8_t*>(flags) + static_cast<size_t>(p) * XW_FLAG_STRIDE);
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (static_cast<int32_t>(*f - epoch) < 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > timeout_ns) {
                atomicOr(dev_status, DEV_ERR_EXCHANGE);
                break;
            }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

// final order of pair_candidates' output over the ranks' runs: type, then key contig in python string order; every (type,
// contig) group comes from ONE rank (the owner of the contig), already in partition / label order, so a row's position is
// its index in its own run plus the rows of the other runs with a smaller key
struct ResRuns {
    const svb_row* rows[EXCH_MAX_WORLD];
    uint32_t n[EXCH_MAX_WORLD], base[EXCH_MAX_WORLD];
    int world;
    uint32_t total;
};
__device__ __forceinline__ unsigned long long order_key(const svb_row& r, const int32_t* lexrank, int n_contig) {
    const int32_t tid = key_contig(r);
    const uint32_t rank = (tid >= 0 && tid < n_contig) ? static_cast<uint32_t>(lexrank[tid]) : 0xFFFFFFFFu;
    return (static_cast<unsigned long long>(r.type) << 32) | rank;
}
__global__ void order_runs_kernel(const __grid_constant__ ResRuns rs, const int32_t* __restrict__ lexrank, int n_contig, svb_row* __restrict__ out) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= rs.total) return;
    int r = 0;
    while (r + 1 < rs.world && x >= rs.base[r + 1]) ++r;
    const uint32_t i = x - rs.base[r];
    svb_row row = rs.rows[r][i];
    const unsigned long long key = order_key(row, lexrank, n_contig);
    uint32_t pos = i;
    for (int q = 0; q < rs.world; ++q) {
        if (q == r) continue;
        uint32_t lo = 0, hi = rs.n[q];
        // rows of run q that come first: smaller key; equal keys cannot occur across runs (one owner per contig) but are
        // ordered by rank for determinism
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const unsigned long long k = order_key(rs.rows[q][mid], lexrank, n_contig);
            if (k < key || (k == key && q < r)) lo = mid + 1; else hi = mid;
        }
        pos += lo;
    }
    row.ordinal = pos;
    out[pos] = row;
}

}  // namespace

struct svb_exchange {
    int device = 0, world = 1, rank = 0;
    uint64_t slot_bytes = 0, res_bytes = 0, window_bytes = 0;
    uint8_t* window = nullptr;                   // own window
    uint8_t* peer[EXCH_MAX_WORLD] = {};          // every rank's window as mapped here (peer[rank] == window)
    bool opened[EXCH_MAX_WORLD] = {};
    unsigned int* d_done = nullptr;              // block counters of the put kernels
    uint64_t* h_headers = nullptr;               // pinned: the slot headers of every rank (8 words each) + the status word, downloaded once per exchange
    uint32_t epoch = 0;
    uint64_t timeout_ns = 20000000000ull;
    uint64_t in_off(int r) const { return static_cast<uint64_t>(r) * slot_bytes; }
    uint64_t res_off(int r) const { return static_cast<uint64_t>(world) * slot_bytes + static_cast<uint64_t>(r) * res_bytes; }
    uint64_t flag_off(int kind, int r) const {
        return static_cast<uint64_t>(world) * (slot_bytes + res_bytes) + (static_cast<uint64_t>(kind) * EXCH_MAX_WORLD + r) * XW_FLAG_STRIDE;
    }
};

static int exchange_put(svb_ctx* ctx, svb_exchange* x, int kind, const PutPiece* pieces, int n_pieces, const uint64_t header[8]) {
    PutArgs a;
    memset(&a, 0, sizeof a);
    a.world = x->world;
    a.n_pieces = n_pieces;
    a.epoch = x->epoch;
    uint64_t total = 0;
    for (int k = 0; k < n_pieces; ++k) { a.piece[k] = pieces[k]; total += pieces[k].bytes; }
    for (int k = 0; k < 8; ++k) a.header[k] = header[k];
    for (int p = 0; p < x->world; ++p) {
        a.peer_slot[p] = x->peer[p] + (kind == 0 ? x->in_off(x->rank) : x->res_off(x->rank));
        a.peer_flag[p] = reinterpret_cast<uint32_t*>(x->peer[p] + x->flag_off(kind, x->rank));
    }
    a.done = x->d_done + kind;
    SVB_CUDA(ctx, cudaMemsetAsync(a.done, 0, sizeof(unsigned int), ctx->stream));
    const unsigned blocks = static_cast<unsigned>(std::min<uint64_t>(std::max<uint64_t>(total / (256 * 64), 1), static_cast<uint64_t>(ctx->sm_count)));
    exchange_put_kernel<<<blocks, 256, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    SVB_CUDA(ctx, cudaGetLastError());
    return SVB_OK;
}

static int exchange_wait(svb_ctx* ctx, svb_exchange* x, int kind) {
    exchange_wait_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(x->window + x->flag_off(kind, 0)), x->world, x->epoch,
                                                    x->timeout_ns, ctx->d_status);
    ctx->launches += 1;
    SVB_CUDA(ctx, cudaGetLastError());
    return SVB_OK;
}

extern "C" {

int svb_exchange_create(svb_ctx* ctx, int world, int rank, uint64_t slot_bytes, uint64_t result_bytes, svb_exchange** out) {
    if (!ctx || !out || world < 1 || world > EXCH_MAX_WORLD || rank < 0 || rank >= world)
        return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_create") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    svb_exchange* x = new (std::nothrow) svb_exchange();
    if (!x) return svb_fail(ctx, SVB_ERR_NOMEM, "svb_exchange_create");
    x->device = ctx->device;
    x->world = world;
    x->rank = rank;
    x->slot_bytes = (std::max<uint64_t>(slot_bytes, 4096) + 255) & ~255ull;
    x->res_bytes = (std::max<uint64_t>(result_bytes, 4096) + 255) & ~255ull;
    x->window_bytes = x->flag_off(2, 0);
    cudaError_t e = cudaMalloc(&x->window, x->window_bytes);          // cudaMalloc, not the stream-ordered pool: IPC needs it
    if (e == cudaSuccess) e = cudaMemset(x->window, 0, x->window_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&x->d_done, 4 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaHostAlloc(&x->h_headers, sizeof(uint64_t) * (8 * EXCH_MAX_WORLD + 1), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        svb_exchange_destroy(ctx, x);
        return svb_fail(ctx, SVB_ERR_CUDA, "svb_exchange_create", e);
    }
    x->peer[rank] = x->window;
    if (const char* env = getenv("SVB_EXCHANGE_TIMEOUT_S")) {
        const long long v = atoll(env);
        if (v > 0) x->timeout_ns = static_cast<uint64_t>(v) * 1000000000ull;
    }
    *out = x;
    return SVB_OK;
}

int svb_exchange_handle(svb_ctx* ctx, const svb_exchange* x, uint8_t handle_out[64]) {
    if (!ctx || !x || !handle_out) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_handle") : SVB_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaSetDevice(ctx->device);
    cudaIpcMemHandle_t h;
    SVB_CUDA(ctx, cudaIpcGetMemHandle(&h, x->window));
    memcpy(handle_out, &h, 64);
    return SVB_OK;
}

int svb_exchange_open(svb_ctx* ctx, svb_exchange* x, const uint8_t* handles) {
    if (!ctx || !x || !handles) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_open") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    for (int p = 0; p < x->world; ++p) {
        if (p == x->rank || x->opened[p]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * static_cast<size_t>(p), 64);
        void* ptr = nullptr;
        SVB_CUDA(ctx, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        x->peer[p] = static_cast<uint8_t*>(ptr);
        x->opened[p] = true;
    }
    return SVB_OK;
}

void svb_exchange_destroy(svb_ctx* ctx, svb_exchange* x) {
    if (!x) return;
    cudaSetDevice(x->device);
    if (ctx && ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int p = 0; p < x->world; ++p)
        if (x->opened[p] && x->peer[p]) cudaIpcCloseMemHandle(x->peer[p]);
    if (x->window) cudaFree(x->window);
    if (x->d_done) cudaFree(x->d_done);
    if (x->h_headers) cudaFreeHost(x->h_headers);
    delete x;
}

// All-gatherv of both haplotype tables over the windows, then "all ranks' rows in append order, restricted to the key
// contigs this rank owns" (svb_exchange_unpack) for both haplotypes.  Collective: every rank calls it once per step.
int svb_exchange_share(svb_ctx* ctx, svb_exchange* x, const svb_table* t1, const svb_table* t2, const int32_t* owner, int n_contig,
                       svb_table** u1, svb_table** u2) {
    if (!ctx || !x || !t1 || !t2 || !owner || !u1 || !u2) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_share") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    *u1 = *u2 = nullptr;
    x->epoch += 1;
    uint64_t sizes[4];
    svb_exchange_sizes(t1, t2, sizes);
    const Layout l = layout_of(sizes);
    if (XW_HEADER + l.total > x->slot_bytes) return svb_fail(ctx, SVB_ERR_CAPACITY, "svb_exchange_share: tables larger than the window slot");
    const svb_table* t[2] = {t1, t2};
    for (int h = 0; h < 2; ++h)
        if (t[h]->n && !t[h]->d_pool_off) return svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_share: gather the tables' sequences first (svb_table_gather_sequences)");
    PutPiece pieces[XW_MAX_PIECES];
    int np = 0;
    for (int h = 0; h < 2; ++h) {
        if (t[h]->n) pieces[np++] = PutPiece{reinterpret_cast<const uint8_t*>(t[h]->d_rows), sizeof(svb_row) * t[h]->n, XW_HEADER + l.rows[h]};
        if (t[h]->d_pool_off && t[h]->n) pieces[np++] = PutPiece{reinterpret_cast<const uint8_t*>(t[h]->d_pool_off), sizeof(uint64_t) * t[h]->n, XW_HEADER + l.off[h]};
        if (t[h]->d_pool_off && t[h]->pool_bytes) pieces[np++] = PutPiece{t[h]->d_pool, t[h]->pool_bytes, XW_HEADER + l.pool[h]};
    }
    uint64_t header[8] = {sizes[0], sizes[1], sizes[2], sizes[3], l.total, x->epoch, 0, 0};
    int rc = exchange_put(ctx, x, 0, pieces, np, header);
    if (rc == SVB_OK) rc = exchange_wait(ctx, x, 0);
    if (rc != SVB_OK) return rc;
    // the sizes of every rank's tables: 64 bytes per slot header, one strided copy, one synchronisation
    uint64_t* const headers = x->h_headers;          // pinned: the two copies are asynchronous, one wait
    SVB_CUDA(ctx, cudaMemcpy2DAsync(headers, XW_HEADER, x->window, x->slot_bytes, XW_HEADER, x->world, cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(headers + 8 * EXCH_MAX_WORLD, ctx->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t h_status = static_cast<uint32_t>(headers[8 * EXCH_MAX_WORLD]);
    if (h_status & DEV_ERR_EXCHANGE) {
        cudaMemsetAsync(ctx->d_status, 0, sizeof(uint32_t), ctx->stream);
        return svb_fail(ctx, SVB_ERR_CUDA, "svb_exchange_share: a peer did not deliver its tables in time");
    }
    std::vector<uint64_t> all_sizes(static_cast<size_t>(x->world) * 4);
    for (int r = 0; r < x->world; ++r) {
        for (int k = 0; k < 4; ++k) all_sizes[4 * r + k] = headers[8 * r + k];
        if (headers[8 * r + 5] != x->epoch) return svb_fail(ctx, SVB_ERR_CUDA, "svb_exchange_share: ranks are out of step");
    }
    // tables that did not bring a pool pack no offsets: unpack reads zeros there (the slot was cleared at creation and pools
    // never shrink to "absent" between steps of one run)
    rc = exchange_unpack_enqueue(ctx, x->window + XW_HEADER, x->slot_bytes, all_sizes.data(), x->world, 1, owner, n_contig, x->rank, u1);
    if (rc == SVB_OK) rc = exchange_unpack_enqueue(ctx, x->window + XW_HEADER, x->slot_bytes, all_sizes.data(), x->world, 2, owner, n_contig, x->rank, u2);
    if (rc == SVB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = svb_fail(ctx, SVB_ERR_CUDA, "svb_exchange_share: unpack");
    if (rc != SVB_OK) {
        if (*u1) { svb_table_free(*u1); *u1 = nullptr; }
        if (*u2) { svb_table_free(*u2); *u2 = nullptr; }
        return rc;
    }
    (*u1)->n = ctx->h_pinned[13];
    (*u2)->n = ctx->h_pinned[14];
    return SVB_OK;
}

// All-gatherv of the paired rows of every rank, put into pair_candidates' order (type, contig by python string order).
// Collective.  contig_lexrank: host array, rank of every contig name.
int svb_exchange_gather_paired(svb_ctx* ctx, svb_exchange* x, const svb_table* paired, const int32_t* contig_lexrank, int n_contig,
                               svb_table** out) {
    if (!ctx || !x || !paired || !out || (n_contig && !contig_lexrank)) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_gather_paired") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    if (XW_HEADER + sizeof(svb_row) * paired->n > x->res_bytes) return svb_fail(ctx, SVB_ERR_CAPACITY, "svb_exchange_gather_paired: rows exceed the result slot");
    PutPiece piece{reinterpret_cast<const uint8_t*>(paired->d_rows), sizeof(svb_row) * paired->n, XW_HEADER};
    uint64_t header[8] = {paired->n, x->epoch, 0, 0, 0, 0, 0, 0};
    int rc = exchange_put(ctx, x, 1, &piece, paired->n ? 1 : 0, header);
    if (rc == SVB_OK) rc = exchange_wait(ctx, x, 1);
    if (rc != SVB_OK) return rc;
    uint64_t* const headers = x->h_headers;
    SVB_CUDA(ctx, cudaMemcpy2DAsync(headers, XW_HEADER, x->window + x->res_off(0), x->res_bytes, XW_HEADER, x->world, cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(headers + 8 * EXCH_MAX_WORLD, ctx->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t h_status = static_cast<uint32_t>(headers[8 * EXCH_MAX_WORLD]);
    if (h_status & DEV_ERR_EXCHANGE) {
        cudaMemsetAsync(ctx->d_status, 0, sizeof(uint32_t), ctx->stream);
        return svb_fail(ctx, SVB_ERR_CUDA, "svb_exchange_gather_paired: a peer did not deliver its rows in time");
    }
    ResRuns rs;
    memset(&rs, 0, sizeof rs);
    rs.world = x->world;
    uint64_t total = 0;
    for (int r = 0; r < x->world; ++r) {
        if (headers[8 * r + 1] != x->epoch) return svb_fail(ctx, SVB_ERR_CUDA, "svb_exchange_gather_paired: ranks are out of step");
        rs.rows[r] = reinterpret_cast<const svb_row*>(x->window + x->res_off(r) + XW_HEADER);
        rs.n[r] = static_cast<uint32_t>(headers[8 * r]);
        rs.base[r] = static_cast<uint32_t>(total);
        total += headers[8 * r];
    }
    if (total >= 0xFFFFFFF0ull) return svb_fail(ctx, SVB_ERR_ARG, "svb_exchange_gather_paired: too many rows");
    rs.total = static_cast<uint32_t>(total);
    svb_table* t = new (std::nothrow) svb_table();
    if (!t) return svb_fail(ctx, SVB_ERR_NOMEM, "svb_exchange_gather_paired");
    t->device = ctx->device;
    t->stream = ctx->stream;
    t->cap = std::max<uint64_t>(total, 1);
    t->n = total;
    int32_t* d_lex = nullptr;
    cudaError_t e = cudaMallocAsync(&t->d_rows, sizeof(svb_row) * t->cap, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_lex, sizeof(int32_t) * std::max(n_contig, 1), ctx->stream);
    if (e == cudaSuccess && n_contig) e = cudaMemcpyAsync(d_lex, contig_lexrank, sizeof(int32_t) * n_contig, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && total) {
        order_runs_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, ctx->stream>>>(rs, d_lex, n_contig, t->d_rows);
        ctx->launches += 1;
        e = cudaGetLastError();
    }
    if (d_lex) cudaFreeAsync(d_lex, ctx->stream);
    if (e != cudaSuccess) {
        svb_table_free(t);
        return svb_fail(ctx, SVB_ERR_CUDA, "svb_exchange_gather_paired", e);
    }
    *out = t;
    return SVB_OK;
}

}  // extern "C"
