// K4 segment_walk + K5 ordered merge.
//
// K4: one thread per primary record that carries SA segments runs walk_read() (walk.cuh), i.e.
// analyze_read_segments (reference SVIM_inter.py:62-340) on the segments that survive the filter
// of analyze_alignment_file_coordsorted (SVIM_COLLECT.py:71-77).  The primary's own coordinates
// (reference_end, query_alignment_start/end, infer_read_length -- pysam properties)
// come from the per-alignment sums that cigar_scan left in HBM plus the clip ops at the two ends
// of its CIGAR run.  Two launches: count, exclusive scan, write -- so rows land in primary order.
//
// K5: the indel table (K2) and the walk table (K4) are each sorted by `ordinal`; a merge-path
// style kernel (one binary search per row) interleaves them into the reference's append order:
// per record, its indels first, then its walk candidates (SVIM_COLLECT.py:79-80).
#include <new>

#include "common.cuh"
#include "walk.cuh"
#include "pairing.cuh"

namespace {

struct WalkArgs {
    const svb_aln_hdr* hdr;
    const uint32_t* cigar;        // flat ops
    const svb_segment* seg;
    const uint32_t* sa_count;
    const uint4* aln_sum;
    const uint32_t* prim_list;
    uint32_t n_prim;
    const int32_t* contig_len;
    const int32_t* contig_lexrank;
    int32_t n_contig;
    WalkParams p;
    uint32_t hap;
    WalkScratch* scratch;
    uint32_t* counts;             // [n_prim + 1]  pass 1: counts; after the scan: exclusive offsets
    svb_row* rows;                // pass 2 destination
    svb_row* stage;               // [n_prim][WALK_STAGE_ROWS] rows kept by the count pass
    uint32_t* dev_status;
    unsigned long long* totals;   // [0] rows of the walk, [1] 4-bit bytes of their inserted sequences (count pass)
    unsigned int* done;           // CTAs of the count pass that have finished
};

constexpr int WALK_THREADS = 32;          // one warp per CTA ...
constexpr int WALK_PER_CTA = 8;           // ... of which 8 lanes carry a primary: the walk is serial, divergent code (ncu: 8 of 32
                                          // lanes active on average, 13 cycles per issued instruction), so its duration is the
                                          // number of DIFFERENT paths a warp has to serialise; few primaries per warp and many
                                          // CTAs (one per SM for a human assembly) shorten it, idle lanes cost nothing
constexpr int WALK_SMEM_SLOTS = 8;        // segments (primary + SA entries) whose scratch lives in shared memory
constexpr uint32_t WALK_STAGE_ROWS = 8;   // rows per primary the count pass keeps, so that the write pass only moves them

// The walk is a chain of dependent reads and writes of a read's scratch slots (sort, three lists, their passes): in global
// memory every step costs an L2 round trip (31 us per pass measured for 850 primaries); reads with up to WALK_SMEM_SLOTS
// segments -- nearly all -- keep them in shared memory, the rest falls back to the global slots.
// WRITE = false: count pass.  Also sums the 4-bit bytes of the inserted sequences (they size the table's sequence pool
// without another round trip to the host), and the LAST CTA to finish turns the counts into exclusive offsets and
// publishes the totals (no separate scan launch).
template <bool WRITE>
__global__ void __launch_bounds__(WALK_THREADS) walk_kernel(const WalkArgs a) {
    __shared__ WalkScratch s_sc[WALK_PER_CTA * WALK_SMEM_SLOTS];
    const uint32_t pi = blockIdx.x * WALK_PER_CTA + threadIdx.x;
    WalkOut o;
    o.rows = nullptr;
    o.cap = 0xFFFFFFFFu;
    o.n = 0;
    o.err = 0;
    o.ins_bytes = 0;
    if (threadIdx.x < static_cast<unsigned>(WALK_PER_CTA) && pi < a.n_prim) {
    if (WRITE) {
        // the count pass left up to WALK_STAGE_ROWS finished rows per primary: nearly every primary only moves them
        const uint32_t first = a.counts[pi], c = a.counts[pi + 1] - first;
        if (c <= WALK_STAGE_ROWS) {
            const uint4* src = reinterpret_cast<const uint4*>(a.stage + static_cast<size_t>(pi) * WALK_STAGE_ROWS);
            uint4* dst = reinterpret_cast<uint4*>(a.rows + first);
            for (uint32_t q = 0; q < 4u * c; ++q) dst[q] = src[q];
            return;
        }
    }
    const uint32_t aln = a.prim_list[pi];
    const svb_aln_hdr h = a.hdr[aln];
    o.rows = WRITE ? a.rows + a.counts[pi] : a.stage + static_cast<size_t>(pi) * WALK_STAGE_ROWS;
    if (!WRITE) o.cap = WALK_STAGE_ROWS;
    const uint4 sums = a.aln_sum[aln];
    // SVIM_COLLECT.py:71 (filter), :73 (supplementary records get no walk), :11-12 (hard clips -> no SA)
    const bool eligible = !(h.flag & 0x4) && !(h.flag & 0x100) && static_cast<int32_t>(h.mapq) >= a.p.min_mapq &&
                          !(h.flag & 0x800) && sums.w == 0u && h.n_cigar > 0u;
    if (eligible) {
        const uint32_t n_sa = a.sa_count[aln];
        WalkScratch* sc = (n_sa + 1u <= static_cast<uint32_t>(WALK_SMEM_SLOTS)) ? s_sc + threadIdx.x * WALK_SMEM_SLOTS
                                                                                 : a.scratch + (static_cast<size_t>(h.sa_first) + pi);
        const uint32_t* ops = a.cigar + h.cigar_off;
        // query_alignment_start: leading S (H skipped)                       [pysam getQueryStart]
        long long qas = 0;
        for (uint32_t k = 0; k < h.n_cigar; ++k) {
            const uint32_t x = ops[k], op = x & 15u;
            if (op == 5u) continue;
            if (op == 4u) qas += x >> 4; else break;
        }
        // query_alignment_end                                                  [pysam getQueryEnd]
        long long qae = h.l_seq;
        if (qae == 0) {
            for (uint32_t k = 0; k < h.n_cigar; ++k) {
                const uint32_t x = ops[k], op = x & 15u;
                if (op == 0u || op == 1u || op == 7u || op == 8u || (op == 4u && qae == 0)) qae += x >> 4;
            }
        } else {
            for (uint32_t k = h.n_cigar - 1u; k >= 1u; --k) {
                const uint32_t x = ops[k], op = x & 15u;
                if (op == 5u) continue;
                if (op == 4u) qae -= x >> 4; else break;
            }
        }
        const long long ref_span = static_cast<long long>(sums.x) + sums.z;            // M,D,=,X + N   [bam_endpos]
        const long long read_len = static_cast<long long>(sums.y) + sums.w;            // M,I,S,=,X + H [infer_read_length]
        WalkSeg prim;
        const bool rev = (h.flag & 0x10) != 0;
        prim.q_start = static_cast<int32_t>(rev ? read_len - qae : qas);               // SVIM_inter.py:70-75
        prim.q_end = static_cast<int32_t>(rev ? read_len - qas : qae);
        prim.tid = h.tid;
        prim.ref_start = h.pos;
        prim.ref_end = static_cast<int32_t>(h.pos + (ref_span ? ref_span : 1));
        prim.rev = rev;
        sc[0].seg = prim;
        uint32_t k = 1;
        for (uint32_t s = 0; s < n_sa; ++s) {
            const svb_segment g = a.seg[h.sa_first + s];
            if (static_cast<int32_t>(g.mapq) < a.p.min_mapq) continue;                 // SVIM_COLLECT.py:77
            WalkSeg w;
            w.q_start = g.is_reverse ? g.read_len - g.q_aend : g.q_astart;
            w.q_end = g.is_reverse ? g.read_len - g.q_astart : g.q_aend;
            w.tid = g.tid;
            w.ref_start = g.pos;
            w.ref_end = g.ref_end;
            w.rev = g.is_reverse;
            sc[k++].seg = w;
        }
        WalkRead rd;
        rd.aln_idx = aln;
        rd.hap = a.hap;
        rd.read_len = static_cast<int32_t>(read_len);
        rd.l_seq = h.l_seq;
        rd.contig_len = a.contig_len;
        rd.contig_lexrank = a.contig_lexrank;
        rd.n_contig = a.n_contig;
        if (k > 1) {
            if (h.tid < 0 || h.tid >= a.n_contig) o.err |= WALK_ERR_BAD_TID;
            else walk_read(rd, a.p, sc, k, o);
        }
    }
    if (!WRITE) a.counts[pi] = o.n;
    if (o.err) {
        uint32_t st = 0;
        if (o.err & WALK_ERR_BAD_TID) st |= DEV_ERR_BAD_TID;
        if (o.err & (WALK_ERR_ASSERT | WALK_ERR_NOSEQ)) st |= DEV_ERR_ASSERT;
        if (o.err & WALK_ERR_CAPACITY) st |= DEV_ERR_CAPACITY;
        atomicOr(a.dev_status, st);
    }
    }
    if (WRITE) return;
    // ---- count pass only: inserted bytes, then the last CTA scans the counts
    const uint32_t bytes = __reduce_add_sync(0xffffffffu, o.ins_bytes);
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        if (bytes) atomicAdd(a.totals + 1, static_cast<unsigned long long>(bytes));
        __threadfence();
        s_last = atomicAdd(a.done, 1u) == gridDim.x - 1u;
    }
    __syncwarp();
    if (!s_last) return;
    __threadfence();
    uint32_t carry = 0;
    for (uint32_t base = 0; base < a.n_prim; base += 32u) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t x = i < a.n_prim ? __ldcg(a.counts + i) : 0u;
        uint32_t inc = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
            if (static_cast<int>(threadIdx.x) >= d) inc += y;
        }
        if (i < a.n_prim) a.counts[i] = carry + inc - x;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (threadIdx.x == 0) {
        a.counts[a.n_prim] = carry;
        a.totals[0] = carry;
    }
}

// exclusive scan of n (+1 total) uint32 in one CTA
__global__ void __launch_bounds__(1024) scan_u32_kernel(uint32_t* v, uint32_t n, unsigned long long* total) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024u) {
        const uint32_t i = base + tid;
        const uint32_t x = i < n ? v[i] : 0u;
        uint32_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (static_cast<int>(lane) >= o) inc += y;
        }
        if (lane == 31u) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, winc, o);
                if (static_cast<int>(lane) >= o) winc += y;
            }
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        const uint32_t excl = s_carry + s_warp[warp] + inc - x;
        if (i < n) v[i] = excl;
        __syncthreads();
        if (tid == 1023u) s_carry = excl + x;
        __syncthreads();
    }
    if (tid == 0) {
        v[n] = s_carry;
        *total = s_carry;
    }
}

__global__ void merge_kernel(const svb_row* __restrict__ A, uint64_t na, const svb_row* __restrict__ B, uint64_t nb,
                             svb_row* __restrict__ out) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= na + nb) return;
    const bool from_a = i < na;
    const svb_row* self = from_a ? A + i : B + (i - na);
    const svb_row* other = from_a ? B : A;
    const uint64_t n_other = from_a ? nb : na;
    const unsigned long long key = self->ordinal;
    uint64_t lo = 0, hi = n_other;                     // ordinals are unique across the two tables
    while (lo < hi) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (other[mid].ordinal < key) lo = mid + 1; else hi = mid;
    }
    const uint64_t dst = (from_a ? i : i - na) + lo;
    const uint4* s = reinterpret_cast<const uint4*>(self);
    uint4* d = reinterpret_cast<uint4*>(out + dst);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}

}  // namespace

// The walk runs in two passes (count, then write at the scanned offsets).  The count pass is enqueued without waiting
// for anything, so that svb_collect reads the indel count, the walk count and the device status back with ONE
// synchronisation; the write pass follows once the host knows how many rows to allocate.
struct WalkPending {
    WalkArgs a;
    unsigned char* base = nullptr;
    unsigned blocks = 0;
};

// `totals`: two device words that receive the number of walk rows and the 4-bit bytes of their inserted sequences
int walk_count_async(svb_ctx* ctx, const svb_records* rec, const svb_params* p, int hap, unsigned long long* totals, WalkPending** out) {
    *out = nullptr;
    SVB_CUDA(ctx, cudaMemsetAsync(totals, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (rec->n_prim == 0) return SVB_OK;
    const size_t scratch_entries = static_cast<size_t>(rec->n_seg) + rec->n_prim;
    const size_t counts_bytes = (static_cast<size_t>(rec->n_prim) + 2) * sizeof(uint32_t) + 16;
    const size_t counts_pad = (counts_bytes + 255) & ~static_cast<size_t>(255);
    WalkPending* w = new (std::nothrow) WalkPending();
    if (!w) return svb_fail(ctx, SVB_ERR_NOMEM, "segment_walk");
    // own allocation (the shared scratch is in use by cigar_scan's unit states on the same stream)
    const size_t scratch_bytes = (scratch_entries * sizeof(WalkScratch) + 255) & ~static_cast<size_t>(255);
    cudaError_t e = cudaMallocAsync(&w->base, counts_pad + scratch_bytes + static_cast<size_t>(rec->n_prim) * WALK_STAGE_ROWS * sizeof(svb_row), ctx->stream);
    if (e != cudaSuccess) {
        delete w;
        return svb_fail(ctx, SVB_ERR_NOMEM, "segment_walk scratch", e);
    }
    WalkArgs& a = w->a;
    a.hdr = rec->d_hdr;
    a.cigar = reinterpret_cast<const uint32_t*>(rec->d_cigar);
    a.seg = rec->d_seg;
    a.sa_count = rec->d_sa_count;
    a.aln_sum = rec->d_aln_sum;
    a.prim_list = rec->d_prim_list;
    a.n_prim = rec->n_prim;
    a.contig_len = rec->d_contig_len;
    a.contig_lexrank = rec->d_contig_lexrank;
    a.n_contig = rec->n_contig;
    a.p.min_mapq = p->min_mapq;
    a.p.min_sv = p->min_sv_size;
    a.p.max_sv = p->max_sv_size;
    a.p.qgt = p->query_gap_tolerance;
    a.p.qot = p->query_overlap_tolerance;
    a.p.rgt = p->reference_gap_tolerance;
    a.p.rot = p->reference_overlap_tolerance;
    a.hap = static_cast<uint32_t>(hap);
    a.counts = reinterpret_cast<uint32_t*>(w->base);
    a.done = reinterpret_cast<unsigned int*>(w->base + (static_cast<size_t>(rec->n_prim) + 2) * sizeof(uint32_t));
    a.scratch = reinterpret_cast<WalkScratch*>(w->base + counts_pad);
    a.stage = reinterpret_cast<svb_row*>(w->base + counts_pad + scratch_bytes);
    a.rows = nullptr;
    a.dev_status = ctx->d_status;
    a.totals = totals;
    w->blocks = (rec->n_prim + WALK_PER_CTA - 1u) / WALK_PER_CTA;
    e = cudaMemsetAsync(a.done, 0, sizeof(unsigned int), ctx->stream);
    if (e == cudaSuccess) {
        KernelTimer timer(ctx, SVB_K_SEGMENT_WALK);
        walk_kernel<false><<<w->blocks, WALK_THREADS, 0, ctx->stream>>>(a);       // count + scan + totals in one launch
        ctx->launches += 1;
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        cudaFreeAsync(w->base, ctx->stream);
        delete w;
        return svb_fail(ctx, SVB_ERR_CUDA, "segment_walk count pass", e);
    }
    *out = w;
    return SVB_OK;
}

void walk_discard(svb_ctx* ctx, WalkPending* w) {
    if (!w) return;
    cudaFreeAsync(w->base, ctx->stream);
    delete w;
}

// n_rows: the count pass's total, read back by the caller.  Consumes `w`.
int walk_write_async(svb_ctx* ctx, WalkPending* w, uint64_t n_rows, svb_row** d_rows_out) {
    *d_rows_out = nullptr;
    if (!w) return SVB_OK;
    int rc = SVB_OK;
    if (n_rows) {
        svb_row* rows = nullptr;
        cudaError_t e = cudaMallocAsync(&rows, sizeof(svb_row) * n_rows, ctx->stream);
        if (e != cudaSuccess) {
            rc = svb_fail(ctx, SVB_ERR_NOMEM, "segment_walk rows", e);
        } else {
            w->a.rows = rows;
            {
                KernelTimer timer(ctx, SVB_K_SEGMENT_WALK);
                walk_kernel<true><<<w->blocks, WALK_THREADS, 0, ctx->stream>>>(w->a);
                ctx->launches += 1;
            }
            e = cudaGetLastError();
            if (e != cudaSuccess) {
                cudaFreeAsync(rows, ctx->stream);
                rc = svb_fail(ctx, SVB_ERR_CUDA, "segment_walk write pass", e);
            } else {
                *d_rows_out = rows;
            }
        }
    }
    walk_discard(ctx, w);
    return rc;
}

int launch_scan_u32(svb_ctx* ctx, uint32_t* v, uint32_t n, unsigned long long* d_total) {
    scan_u32_kernel<<<1, 1024, 0, ctx->stream>>>(v, n, d_total);
    ctx->launches += 1;
    SVB_CUDA(ctx, cudaGetLastError());
    return SVB_OK;
}

int launch_merge_tables(svb_ctx* ctx, const svb_row* a, uint64_t na, const svb_row* b, uint64_t nb, svb_row* out) {
    const uint64_t n = na + nb;
    if (!n) return SVB_OK;
    KernelTimer timer(ctx, SVB_K_MERGE);
    merge_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(a, na, b, nb, out);
    ctx->launches += 1;
    SVB_CUDA(ctx, cudaGetLastError());
    return SVB_OK;
}
