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
    uint32_t* dev_status;
};

template <bool WRITE>
__global__ void __launch_bounds__(128) walk_kernel(const WalkArgs a) {
    const uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= a.n_prim) return;
    const uint32_t aln = a.prim_list[pi];
    const svb_aln_hdr h = a.hdr[aln];
    WalkOut o;
    o.rows = WRITE ? a.rows + a.counts[pi] : nullptr;
    o.n = 0;
    o.err = 0;
    const uint4 sums = a.aln_sum[aln];
    // SVIM_COLLECT.py:71 (filter), :73 (supplementary records get no walk), :11-12 (hard clips -> no SA)
    const bool eligible = !(h.flag & 0x4) && !(h.flag & 0x100) && static_cast<int32_t>(h.mapq) >= a.p.min_mapq &&
                          !(h.flag & 0x800) && sums.w == 0u && h.n_cigar > 0u;
    if (eligible) {
        WalkScratch* sc = a.scratch + (static_cast<size_t>(h.sa_first) + pi);
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
        const uint32_t n_sa = a.sa_count[aln];
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

int walk_count_async(svb_ctx* ctx, const svb_records* rec, const svb_params* p, int hap, WalkPending** out) {
    *out = nullptr;
    SVB_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 1, 0, sizeof(unsigned long long), ctx->stream));
    if (rec->n_prim == 0) return SVB_OK;
    const size_t scratch_entries = static_cast<size_t>(rec->n_seg) + rec->n_prim;
    const size_t counts_bytes = (static_cast<size_t>(rec->n_prim) + 2) * sizeof(uint32_t);
    const size_t counts_pad = (counts_bytes + 255) & ~static_cast<size_t>(255);
    WalkPending* w = new (std::nothrow) WalkPending();
    if (!w) return svb_fail(ctx, SVB_ERR_NOMEM, "segment_walk");
    // own allocation (the shared scratch is in use by cigar_scan's unit states on the same stream)
    cudaError_t e = cudaMallocAsync(&w->base, counts_pad + scratch_entries * sizeof(WalkScratch), ctx->stream);
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
    a.scratch = reinterpret_cast<WalkScratch*>(w->base + counts_pad);
    a.rows = nullptr;
    a.dev_status = ctx->d_status;
    w->blocks = (rec->n_prim + 127u) / 128u;
    {
        KernelTimer timer(ctx, SVB_K_SEGMENT_WALK);
        walk_kernel<false><<<w->blocks, 128, 0, ctx->stream>>>(a);
        scan_u32_kernel<<<1, 1024, 0, ctx->stream>>>(a.counts, rec->n_prim, ctx->d_counters + 1);
        ctx->launches += 2;
    }
    e = cudaGetLastError();
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
                walk_kernel<true><<<w->blocks, 128, 0, ctx->stream>>>(w->a);
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
