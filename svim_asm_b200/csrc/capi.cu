// C-ABI glue of libsvimasm_b200.so (include/svimasm_b200.h): context, uploads, collect, tables.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"
#include "pairing.cuh"

int svb_fail(svb_ctx* ctx, int code, const char* what, cudaError_t e) {
    if (ctx) {
        char buf[512];
        if (e != cudaSuccess)
            snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
        else
            snprintf(buf, sizeof buf, "%s", what);
        ctx->err = buf;
    }
    return code;
}

KernelTimer::KernelTimer(svb_ctx* c, int kernel_id) : ctx(c), live(false) {
    if (!c->timing_enabled) return;
    auto take = [&](cudaEvent_t* ev) {
        if (!c->free_events.empty()) {
            *ev = c->free_events.back();
            c->free_events.pop_back();
            return true;
        }
        return cudaEventCreate(ev) == cudaSuccess;
    };
    span.kernel = kernel_id;
    if (!take(&span.start)) return;
    if (!take(&span.stop)) {
        c->free_events.push_back(span.start);
        return;
    }
    cudaEventRecord(span.start, c->stream);
    live = true;
}

KernelTimer::~KernelTimer() {
    if (!live) return;
    cudaEventRecord(span.stop, ctx->stream);
    ctx->spans.push_back(span);
}

void* svb_scratch(svb_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return ctx->d_scratch;
    if (ctx->d_scratch) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->d_scratch);
        ctx->d_scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    size_t want = std::max(bytes, static_cast<size_t>(1) << 20);
    want = (want + 0xFFFFF) & ~static_cast<size_t>(0xFFFFF);
    if (cudaMalloc(&ctx->d_scratch, want) != cudaSuccess) return nullptr;
    ctx->scratch_bytes = want;
    return ctx->d_scratch;
}

static int check_device_status(svb_ctx* ctx) {
    // one 4-byte readback; called where the host synchronises anyway
    uint32_t st = 0;
    SVB_CUDA(ctx, cudaMemcpyAsync(&st, ctx->d_status, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!st) return SVB_OK;
    cudaMemsetAsync(ctx->d_status, 0, sizeof(uint32_t), ctx->stream);
    if (st & DEV_ERR_BAD_TID)
        return svb_fail(ctx, SVB_ERR_FORMAT, "reference_id out of range (pysam get_reference_name would raise ValueError)");
    if (st & DEV_ERR_ASSERT)
        return svb_fail(ctx, SVB_ERR_ASSERT, "candidate end is smaller than its start (reference assertion, SVCandidate.py)");
    if (st & DEV_ERR_NOSEQ)
        return svb_fail(ctx, SVB_ERR_ARG, "insertion candidates need the query sequences: call svb_records_set_sequences first");
    if (st & DEV_ERR_EXCHANGE) return svb_fail(ctx, SVB_ERR_CUDA, "multi-GPU exchange: a peer did not arrive in time");
    return svb_fail(ctx, SVB_ERR_CAPACITY, "per-read scratch capacity exceeded");
}

extern "C" {

int svb_abi_version(void) { return SVB_ABI_VERSION; }

int svb_create(int device, svb_ctx** out) {
    if (!out) return SVB_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return SVB_ERR_CUDA;
    svb_ctx* ctx = new (std::nothrow) svb_ctx();
    if (!ctx) return SVB_ERR_NOMEM;
    ctx->device = device;
    memset(&ctx->timing, 0, sizeof ctx->timing);
    if (const char* env = getenv("SVB_ED_STRIDE")) {          // test hook: a small slice exercises the grow-and-repeat path
        const long long v = atoll(env);
        if (v > 0) ctx->ed_stride = static_cast<uint64_t>(v);
    }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    ctx->main_stream = ctx->stream;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_status, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_status, 0, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_counters, 64 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_counters, 0, 64 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_pinned, 64 * sizeof(unsigned long long));
    if (e == cudaSuccess) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;      // keep freed blocks cached: tables are re-allocated every step
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (e != cudaSuccess) {
        svb_destroy(ctx);
        return SVB_ERR_CUDA;
    }
    *out = ctx;
    return SVB_OK;
}

void svb_destroy(svb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& s : ctx->spans) {
        cudaEventDestroy(s.start);
        cudaEventDestroy(s.stop);
    }
    for (auto e : ctx->free_events) cudaEventDestroy(e);
    for (auto e : ctx->marks)
        if (e) cudaEventDestroy(e);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->d_status) cudaFree(ctx->d_status);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_text) cudaFreeHost(ctx->h_text);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    upload_release(ctx);
    if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->main_stream) cudaStreamDestroy(ctx->main_stream);
    delete ctx;
}

const char* svb_last_error(const svb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int svb_synchronize(svb_ctx* ctx) {
    if (!ctx) return SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SVB_OK;
}

static int fold_spans(svb_ctx* ctx) {
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& s : ctx->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.start, s.stop) == cudaSuccess && s.kernel >= 0 && s.kernel < SVB_K_COUNT) {
            ctx->timing.ms[s.kernel] += ms;
            ctx->timing.launches[s.kernel] += 1;
        }
        ctx->free_events.push_back(s.start);
        ctx->free_events.push_back(s.stop);
    }
    ctx->spans.clear();
    return SVB_OK;
}

int svb_timing_reset(svb_ctx* ctx) {
    if (!ctx) return SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    int rc = fold_spans(ctx);
    memset(&ctx->timing, 0, sizeof ctx->timing);
    return rc;
}

int svb_timing_get(svb_ctx* ctx, svb_timing* out) {
    if (!ctx || !out) return SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    int rc = fold_spans(ctx);
    *out = ctx->timing;
    return rc;
}

int svb_mark(svb_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 16) return SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (!ctx->marks[slot]) SVB_CUDA(ctx, cudaEventCreate(&ctx->marks[slot]));
    SVB_CUDA(ctx, cudaEventRecord(ctx->marks[slot], ctx->stream));
    return SVB_OK;
}

int svb_elapsed_ms(svb_ctx* ctx, int slot_begin, int slot_end, double* ms) {
    if (!ctx || !ms || slot_begin < 0 || slot_begin >= 16 || slot_end < 0 || slot_end >= 16 || !ctx->marks[slot_begin] ||
        !ctx->marks[slot_end])
        return SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    SVB_CUDA(ctx, cudaEventSynchronize(ctx->marks[slot_end]));
    float f = 0.f;
    SVB_CUDA(ctx, cudaEventElapsedTime(&f, ctx->marks[slot_begin], ctx->marks[slot_end]));
    *ms = f;
    return SVB_OK;
}

int svb_launch_count(svb_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return SVB_ERR_ARG;
    *out = ctx->launches;
    return SVB_OK;
}

int svb_pair_stats(svb_ctx* ctx, uint64_t out[4]) {
    if (!ctx || !out) return SVB_ERR_ARG;
    for (int i = 0; i < 4; ++i) out[i] = ctx->last_pair_stats[i];
    return SVB_OK;
}

int svb_set_scan_variant(svb_ctx* ctx, int variant) {
    if (!ctx || variant < 0 || variant > 1) return SVB_ERR_ARG;
    ctx->scan_variant = variant;
    return SVB_OK;
}

// ---- records ---------------------------------------------------------------------------------

static void free_async(void* p, cudaStream_t st) {
    if (p) cudaFreeAsync(p, st);
}

void svb_records_free(svb_records* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    if (r->seq_borrowed) { r->d_seq4 = nullptr; r->d_seq_off = nullptr; }      // pinned host memory of the caller, not ours
    void* ptrs[] = {r->d_hdr, r->d_cigar, r->d_off4, r->d_chunk_first, r->d_seg, r->d_sa_count, r->d_contig_len,
                    r->d_contig_lexrank, r->d_aln_sum, r->d_prim_list, r->d_seq4, r->d_seq_off, r->d_global_idx};
    for (void* q : ptrs) free_async(q, r->stream);
    delete r;
}

// `d_cigar_prebuilt`: the CIGAR array already on the device (allocated with cigar_padded_n4 uint4; the device ingest
// builds it there); ownership passes to the records.  Otherwise `cigar` is a host array that is uploaded.
int load_records_impl(svb_ctx* ctx, const svb_aln_hdr* hdr, uint32_t n_aln, const uint32_t* cigar, uint4* d_cigar_prebuilt,
                      uint64_t n_ops_padded, const svb_segment* seg, const uint32_t* sa_count, uint32_t n_seg,
                      const int32_t* contig_len, const int32_t* contig_lexrank, int32_t n_contig,
                      svb_records** out) {
    if (!ctx || !out || (n_aln && !hdr) || (n_ops_padded && !cigar && !d_cigar_prebuilt) || n_contig < 0 || (n_contig && (!contig_len || !contig_lexrank)))
        return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_load_records: null argument") : SVB_ERR_ARG;
    if (n_ops_padded % 4) return svb_fail(ctx, SVB_ERR_ARG, "svb_load_records: n_ops_padded must be a multiple of 4");
    if (n_ops_padded / 4 >= 0xFFFFFFFFull) return svb_fail(ctx, SVB_ERR_ARG, "svb_load_records: more than 2^34 ops");
    if (n_seg && (!seg || !sa_count)) return svb_fail(ctx, SVB_ERR_ARG, "svb_load_records: segments without counts");
    cudaSetDevice(ctx->device);
    *out = nullptr;
    svb_records* r = new (std::nothrow) svb_records();
    if (!r) return svb_fail(ctx, SVB_ERR_NOMEM, "svb_load_records");
    r->device = ctx->device;
    r->stream = ctx->stream;
    r->n_aln = n_aln;
    r->n_seg = n_seg;
    r->n_contig = n_contig;
    r->n4 = n_ops_padded / 4;
    if (n_contig) r->h_contig_len.assign(contig_len, contig_len + n_contig);

    // host-side derived index: off4[] and the list of records that carry SA segments
    std::vector<uint32_t> off4(static_cast<size_t>(n_aln) + 1), prim;
    uint64_t expect = 0;
    uint64_t seg_seen = 0;
    for (uint32_t i = 0; i < n_aln; ++i) {
        if (hdr[i].cigar_off % 4 || hdr[i].cigar_off != expect) {
            delete r;
            return svb_fail(ctx, SVB_ERR_ARG, "svb_load_records: CIGAR runs must be contiguous, in record order and 4-op aligned");
        }
        off4[i] = static_cast<uint32_t>(hdr[i].cigar_off / 4);
        expect += (static_cast<uint64_t>(hdr[i].n_cigar) + 3) / 4 * 4;
        r->n_ops += hdr[i].n_cigar;
        const uint32_t cnt = sa_count ? sa_count[i] : 0;
        if (cnt) {
            if (hdr[i].sa_first != seg_seen) {
                delete r;
                return svb_fail(ctx, SVB_ERR_ARG, "svb_load_records: sa_first must index seg[] in record order");
            }
            prim.push_back(i);
            seg_seen += cnt;
        }
    }
    if (expect != n_ops_padded || seg_seen != n_seg) {
        delete r;
        return svb_fail(ctx, SVB_ERR_ARG, "svb_load_records: totals do not match the per-record counts");
    }
    off4[n_aln] = static_cast<uint32_t>(r->n4);
    r->n_prim = static_cast<uint32_t>(prim.size());

    auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
        *dst = nullptr;
        if (!bytes) return cudaSuccess;
        cudaError_t e = cudaMallocAsync(dst, bytes, ctx->stream);      // stream-ordered pool: no device-wide sync per step
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    };
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = up(reinterpret_cast<void**>(&r->d_hdr), hdr, sizeof(svb_aln_hdr) * n_aln);
    if (e == cudaSuccess && d_cigar_prebuilt) {
        r->d_cigar = d_cigar_prebuilt;
    } else if (e == cudaSuccess) {   // allocated in whole scan units; launch_build_chunk_index fills the tail with op 15
        e = cudaMallocAsync(reinterpret_cast<void**>(&r->d_cigar), std::max<uint64_t>(cigar_padded_n4(r->n4), 1) * sizeof(uint4), ctx->stream);
        if (e == cudaSuccess && n_ops_padded)
            e = cudaMemcpyAsync(r->d_cigar, cigar, sizeof(uint32_t) * n_ops_padded, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess) e = up(reinterpret_cast<void**>(&r->d_off4), off4.data(), sizeof(uint32_t) * off4.size());
    if (e == cudaSuccess) e = up(reinterpret_cast<void**>(&r->d_seg), seg, sizeof(svb_segment) * n_seg);
    if (e == cudaSuccess) e = up(reinterpret_cast<void**>(&r->d_sa_count), sa_count, sa_count ? sizeof(uint32_t) * n_aln : 0);
    if (e == cudaSuccess) e = up(reinterpret_cast<void**>(&r->d_contig_len), contig_len, sizeof(int32_t) * n_contig);
    if (e == cudaSuccess) e = up(reinterpret_cast<void**>(&r->d_contig_lexrank), contig_lexrank, sizeof(int32_t) * n_contig);
    if (e == cudaSuccess) e = up(reinterpret_cast<void**>(&r->d_prim_list), prim.data(), sizeof(uint32_t) * prim.size());
    if (e == cudaSuccess && n_aln) e = cudaMallocAsync(&r->d_aln_sum, sizeof(uint4) * n_aln, ctx->stream);
    if (e != cudaSuccess) {
        cudaStreamSynchronize(ctx->stream);
        svb_records_free(r);
        return svb_fail(ctx, SVB_ERR_CUDA, "svb_load_records upload", e);
    }
    int rc = launch_build_chunk_index(ctx, r);
    if (rc == SVB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = svb_fail(ctx, SVB_ERR_CUDA, "svb_load_records sync");
    if (rc != SVB_OK) {
        svb_records_free(r);
        return rc;
    }
    *out = r;
    return SVB_OK;
}

int svb_load_records(svb_ctx* ctx, const svb_aln_hdr* hdr, uint32_t n_aln, const uint32_t* cigar,
                     uint64_t n_ops_padded, const svb_segment* seg, const uint32_t* sa_count, uint32_t n_seg,
                     const int32_t* contig_len, const int32_t* contig_lexrank, int32_t n_contig,
                     svb_records** out) {
    return load_records_impl(ctx, hdr, n_aln, cigar, nullptr, n_ops_padded, seg, sa_count, n_seg, contig_len, contig_lexrank, n_contig, out);
}

int svb_records_set_sequences(svb_ctx* ctx, svb_records* rec, const uint8_t* seq4, const uint64_t* seq_off) {
    if (!ctx || !rec || !seq_off) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_records_set_sequences") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    const uint64_t bytes = seq_off[rec->n_aln];
    if (!rec->seq_borrowed) {
        free_async(rec->d_seq4, ctx->stream);
        free_async(rec->d_seq_off, ctx->stream);
    }
    rec->seq_borrowed = false;
    rec->d_seq4 = nullptr;
    rec->d_seq_off = nullptr;
    SVB_CUDA(ctx, cudaMallocAsync(&rec->d_seq_off, sizeof(uint64_t) * (static_cast<size_t>(rec->n_aln) + 1), ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(rec->d_seq_off, seq_off, sizeof(uint64_t) * (static_cast<size_t>(rec->n_aln) + 1),
                                  cudaMemcpyHostToDevice, ctx->stream));
    if (bytes) {
        SVB_CUDA(ctx, cudaMallocAsync(&rec->d_seq4, bytes, ctx->stream));
        SVB_CUDA(ctx, cudaMemcpyAsync(rec->d_seq4, seq4, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    rec->seq_bytes = bytes;
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SVB_OK;
}

// The query sequences stay where they are -- in PINNED host memory of the caller -- and the device reads the few MB of
// inserted bases it needs in place over PCIe (sequence pools of svb_collect2 / svb_table_gather_sequences) instead of the
// 0.65 GB of a whole assembly being uploaded.  The buffers must outlive the records.
int svb_records_map_sequences_host(svb_ctx* ctx, svb_records* rec, const uint8_t* seq4_pinned, const uint64_t* seq_off_pinned) {
    if (!ctx || !rec || !seq_off_pinned) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_records_map_sequences_host") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    auto view = [](const void* p) -> void* {
        cudaPointerAttributes attr;
        if (!p || cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged) ? attr.devicePointer : nullptr;
    };
    void* d_off = view(seq_off_pinned);
    void* d_seq = seq4_pinned ? view(seq4_pinned) : nullptr;
    if (!d_off || (seq4_pinned && !d_seq)) return svb_fail(ctx, SVB_ERR_ARG, "svb_records_map_sequences_host: the buffers are not pinned host memory");
    if (!rec->seq_borrowed) {
        free_async(rec->d_seq4, ctx->stream);
        free_async(rec->d_seq_off, ctx->stream);
    }
    rec->d_seq4 = static_cast<uint8_t*>(d_seq);
    rec->d_seq_off = static_cast<uint64_t*>(d_off);
    rec->seq_bytes = 0;
    rec->seq_borrowed = true;
    return SVB_OK;
}

// ---- tables ----------------------------------------------------------------------------------

static svb_table* table_alloc(svb_ctx* ctx, uint64_t cap) {
    svb_table* t = new (std::nothrow) svb_table();
    if (!t) return nullptr;
    t->device = ctx->device;
    t->stream = ctx->stream;
    t->cap = std::max<uint64_t>(cap, 1);
    if (cudaMallocAsync(&t->d_rows, sizeof(svb_row) * t->cap, ctx->stream) != cudaSuccess) {
        delete t;
        return nullptr;
    }
    return t;
}

void svb_table_free(svb_table* t) {
    if (!t) return;
    cudaSetDevice(t->device);
    free_async(t->d_rows, t->stream);
    table_drop_pool(t);
    delete t;
}

int64_t svb_table_size(const svb_table* t) { return t ? static_cast<int64_t>(t->n) : -1; }

int svb_table_to_host(svb_ctx* ctx, const svb_table* t, svb_row* dst, uint64_t cap, uint64_t* n) {
    if (!ctx || !t || !n) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_table_to_host") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    *n = t->n;
    const uint64_t m = std::min(cap, t->n);
    if (m && !dst) return svb_fail(ctx, SVB_ERR_ARG, "svb_table_to_host: null destination");
    if (m) {
        // a copy into pageable memory is staged by the driver in small pieces (0.15 ms for a whole-genome table): go through
        // the context's pinned buffer instead (one DMA at link speed, then a host memcpy), unless `dst` is pinned itself
        const size_t bytes = sizeof(svb_row) * m;
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, dst) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (!pinned && bytes <= (64u << 20)) {
            if (ctx->h_stage_cap < bytes) {
                if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
                ctx->h_stage = nullptr;
                ctx->h_stage_cap = 0;
                const size_t want = std::max<size_t>(bytes * 2, 1u << 20);
                if (cudaMallocHost(&ctx->h_stage, want) == cudaSuccess) ctx->h_stage_cap = want;
                else cudaGetLastError();
            }
        }
        if (!pinned && ctx->h_stage_cap >= bytes) {
            SVB_CUDA(ctx, cudaMemcpyAsync(ctx->h_stage, t->d_rows, bytes, cudaMemcpyDeviceToHost, ctx->stream));
            SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            memcpy(dst, ctx->h_stage, bytes);
            return SVB_OK;
        }
        SVB_CUDA(ctx, cudaMemcpyAsync(dst, t->d_rows, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SVB_OK;
}

int svb_table_from_host(svb_ctx* ctx, const svb_row* rows, uint64_t n, svb_table** out) {
    if (!ctx || !out || (n && !rows)) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_table_from_host") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    svb_table* t = table_alloc(ctx, n);
    if (!t) return svb_fail(ctx, SVB_ERR_NOMEM, "svb_table_from_host");
    t->n = n;
    if (n) {
        cudaError_t e = cudaMemcpyAsync(t->d_rows, rows, sizeof(svb_row) * n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            svb_table_free(t);
            return svb_fail(ctx, SVB_ERR_CUDA, "svb_table_from_host", e);
        }
    }
    *out = t;
    return SVB_OK;
}

int svb_table_export(svb_ctx* ctx, const svb_table* t, void* device_dst, uint64_t cap_rows) {
    if (!ctx || !t || (t->n && !device_dst)) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_table_export") : SVB_ERR_ARG;
    if (t->n > cap_rows) return svb_fail(ctx, SVB_ERR_ARG, "svb_table_export: destination too small");
    cudaSetDevice(ctx->device);
    if (t->n) SVB_CUDA(ctx, cudaMemcpyAsync(device_dst, t->d_rows, sizeof(svb_row) * t->n, cudaMemcpyDeviceToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SVB_OK;
}

int svb_table_import(svb_ctx* ctx, const void* device_src, uint64_t n_rows, svb_table** out) {
    if (!ctx || !out || (n_rows && !device_src)) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_table_import") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    svb_table* t = table_alloc(ctx, n_rows);
    if (!t) return svb_fail(ctx, SVB_ERR_NOMEM, "svb_table_import");
    t->n = n_rows;
    if (n_rows) {
        cudaError_t e = cudaMemcpyAsync(t->d_rows, device_src, sizeof(svb_row) * n_rows, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            svb_table_free(t);
            return svb_fail(ctx, SVB_ERR_CUDA, "svb_table_import", e);
        }
    }
    *out = t;
    return SVB_OK;
}

// ---- collect -----------------------------------------------------------------------------------

// ---- collect: enqueue (K2 + finalize + K4 count), ONE synchronisation for the counts, finish (K4 write + K5 merge) ----
// A collect is split so that the two haplotypes of a diploid run share the synchronisation (svb_collect2): everything up to
// the counts of both is enqueued, the host waits once, then the walk rows are written and merged without waiting again.
// Device counters of a pending collect (base = 0 or 16): [base] indel rows, [base + 1] walk rows, [base + 2] 4-bit bytes of
// the walk's inserted sequences, [base + 3] those of the indel rows.
namespace {

struct CollectPending {
    const svb_records* rec = nullptr;
    int hap = 0, base = 0;
    svb_table* indel = nullptr;
    WalkPending* walk = nullptr;
};

void collect_drop(svb_ctx* ctx, CollectPending& c) {
    if (c.indel) svb_table_free(c.indel);
    walk_discard(ctx, c.walk);
    c.indel = nullptr;
    c.walk = nullptr;
}

int collect_begin(svb_ctx* ctx, const svb_records* rec, const svb_params* p, int hap, int base, uint64_t cap, bool with_walk, CollectPending* c) {
    c->rec = rec;
    c->hap = hap;
    c->base = base;
    c->indel = table_alloc(ctx, cap);
    if (!c->indel) return svb_fail(ctx, SVB_ERR_NOMEM, "svb_collect: indel table");
    // (the bytes of inserted sequence the rows hold -- the size of the sequence pool -- are summed by the scan's row finalize)
    ScanOutput so{c->indel->d_rows, c->indel->cap, ctx->d_counters + base, ctx->d_counters + base + 3};
    int rc = launch_cigar_scan(ctx, rec, p, hap, so);
    if (rc == SVB_OK && with_walk) rc = walk_count_async(ctx, rec, p, hap, ctx->d_counters + base + 1, &c->walk);
    if (rc != SVB_OK) collect_drop(ctx, *c);
    return rc;
}

// the counts of the pending collects + the device status, one wait
int collect_sync(svb_ctx* ctx, CollectPending* list, int n) {
    uint32_t* h_status = reinterpret_cast<uint32_t*>(ctx->h_pinned + 12);
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < n && e == cudaSuccess; ++k)
        e = cudaMemcpyAsync(ctx->h_pinned + list[k].base, ctx->d_counters + list[k].base, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_status, ctx->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return svb_fail(ctx, SVB_ERR_CUDA, "svb_collect: cigar_scan", e);
    if (*h_status) return check_device_status(ctx);          // the kernels flagged something the reference would have raised on
    return SVB_OK;
}

// after collect_sync: K4 write pass, K5 merge, optionally the sequence pool -- all stream ordered, no further wait
int collect_end(svb_ctx* ctx, CollectPending& c, const svb_params* p, bool with_pool, svb_table** out) {
    *out = nullptr;
    const unsigned long long n_indel = ctx->h_pinned[c.base], n_walk = c.walk ? ctx->h_pinned[c.base + 1] : 0ull;
    const unsigned long long pool_bytes = ctx->h_pinned[c.base + 3] + (c.walk ? ctx->h_pinned[c.base + 2] : 0ull);
    if (n_indel > c.indel->cap) {
        // the capacity guess (1 row per 512 ops) was too small: scan once more with the exact size (the walk's count pass is
        // still valid: it reads the per-alignment sums, not the rows)
        svb_table_free(c.indel);
        c.indel = nullptr;
        CollectPending again;
        int rc = collect_begin(ctx, c.rec, p, c.hap, c.base, n_indel, false, &again);
        if (rc == SVB_OK) {
            again.walk = nullptr;
            rc = collect_sync(ctx, &again, 1);
        }
        if (rc != SVB_OK) {
            collect_drop(ctx, again);
            collect_drop(ctx, c);
            return rc;
        }
        c.indel = again.indel;
        if (ctx->h_pinned[c.base] > c.indel->cap) {
            collect_drop(ctx, c);
            return svb_fail(ctx, SVB_ERR_CAPACITY, "svb_collect: indel table overflow");
        }
    }
    svb_table* indel = c.indel;
    c.indel = nullptr;
    indel->n = ctx->h_pinned[c.base];
    indel->stream = ctx->stream;          // its last readers (merge, pool gather) run on the stream this finish is enqueued on
    // K4: split-alignment walk rows (own table, emission order per primary)
    svb_row* d_walk = nullptr;
    int rc = walk_write_async(ctx, c.walk, n_walk, &d_walk);
    c.walk = nullptr;
    if (rc != SVB_OK) {
        svb_table_free(indel);
        return rc;
    }
    svb_table* result = indel;
    if (n_walk) {
        // K5: order-preserving merge by ordinal (record order; indels of a record before its walk rows)
        svb_table* merged = table_alloc(ctx, indel->n + n_walk);
        if (!merged) {
            svb_table_free(indel);
            free_async(d_walk, ctx->stream);
            return svb_fail(ctx, SVB_ERR_NOMEM, "svb_collect: merged table");
        }
        rc = launch_merge_tables(ctx, indel->d_rows, indel->n, d_walk, n_walk, merged->d_rows);     // stream ordered: no wait
        merged->n = indel->n + n_walk;
        svb_table_free(indel);
        free_async(d_walk, ctx->stream);
        if (rc != SVB_OK) {
            svb_table_free(merged);
            return rc;
        }
        result = merged;
    }
    if (with_pool) {
        if (!c.rec->d_seq_off) {
            svb_table_free(result);
            return svb_fail(ctx, SVB_ERR_ARG, "svb_collect2: sequence pools need the query sequences (svb_records_set_sequences)");
        }
        rc = gather_pool_known(ctx, result, c.rec->d_seq4, c.rec->d_seq_off, pool_bytes);
        if (rc != SVB_OK) {
            svb_table_free(result);
            return rc;
        }
    }
    *out = result;
    return SVB_OK;
}

uint64_t collect_cap_guess(const svb_records* rec) {
    // K2's row capacity is a guess (1 row per 512 ops; human assemblies have about 1 per 20,000); an overflow is detected from
    // the exact count and that scan is repeated once with the right size
    return std::max<uint64_t>(4096, rec->n_ops / 512);
}

}  // namespace

int svb_collect(svb_ctx* ctx, const svb_records* rec, const svb_params* p, int hap, svb_table** out) {
    if (!ctx || !rec || !p || !out || hap < 0 || hap > 2) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_collect") : SVB_ERR_ARG;
    if (rec->device != ctx->device) return svb_fail(ctx, SVB_ERR_ARG, "svb_collect: records live on another device");
    cudaSetDevice(ctx->device);
    *out = nullptr;
    CollectPending c;
    int rc = collect_begin(ctx, rec, p, hap, 0, collect_cap_guess(rec), true, &c);
    if (rc != SVB_OK) return rc;
    rc = collect_sync(ctx, &c, 1);
    if (rc != SVB_OK) {
        collect_drop(ctx, c);
        return rc;
    }
    return collect_end(ctx, c, p, false, out);
}

int svb_collect2(svb_ctx* ctx, const svb_records* rec1, const svb_records* rec2, const svb_params* p, int with_pools, svb_table** out1,
                 svb_table** out2) {
    if (!ctx || !rec1 || !rec2 || !p || !out1 || !out2) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_collect2") : SVB_ERR_ARG;
    if (rec1->device != ctx->device || rec2->device != ctx->device) return svb_fail(ctx, SVB_ERR_ARG, "svb_collect2: records live on another device");
    cudaSetDevice(ctx->device);
    *out1 = *out2 = nullptr;
    // The scans are bandwidth-bound and fill the GPU; the walk is a few dozen latency-bound warps.  Haplotype 1's walk
    // count runs on the side stream while haplotype 2 is scanned, and after the one synchronisation the two finishing chains
    // (walk write, merge, pool) run side by side.
    const bool overlap = !getenv("SVB_COLLECT_ONE_STREAM");
    CollectPending c[2];
    int rc = collect_begin(ctx, rec1, p, 1, 0, collect_cap_guess(rec1), !overlap, &c[0]);
    if (rc != SVB_OK) return rc;
    if (overlap) {
        SideStream side(ctx);
        rc = walk_count_async(ctx, rec1, p, 1, ctx->d_counters + 1, &c[0].walk);
    }
    if (rc == SVB_OK) rc = collect_begin(ctx, rec2, p, 2, 16, collect_cap_guess(rec2), true, &c[1]);
    if (overlap) SideStream::join(ctx);
    if (rc == SVB_OK) rc = collect_sync(ctx, c, 2);
    if (rc != SVB_OK) {
        collect_drop(ctx, c[0]);
        collect_drop(ctx, c[1]);
        return rc;
    }
    int rc2 = SVB_OK;
    if (overlap) {
        {
            SideStream side(ctx);
            rc2 = collect_end(ctx, c[1], p, with_pools != 0, out2);
            if (rc2 == SVB_OK) {            // allocated on the side stream, used and freed on the main one from now on
                (*out2)->stream = ctx->main_stream;
            }
        }
        rc = collect_end(ctx, c[0], p, with_pools != 0, out1);
        SideStream::join(ctx);
    } else {
        rc = collect_end(ctx, c[0], p, with_pools != 0, out1);
        if (rc == SVB_OK) rc2 = collect_end(ctx, c[1], p, with_pools != 0, out2);
    }
    if (rc == SVB_OK) rc = rc2;
    if (rc != SVB_OK) {
        collect_drop(ctx, c[0]);
        collect_drop(ctx, c[1]);
        if (*out1) { svb_table_free(*out1); *out1 = nullptr; }
        if (*out2) { svb_table_free(*out2); *out2 = nullptr; }
    }
    return rc;
}

int svb_cigar_indel(svb_ctx* ctx, const uint32_t* packed_ops, uint32_t n_ops, int32_t min_length, int64_t* out4,
                    uint32_t cap, uint32_t* n_out) {
    if (!ctx || !n_out || (n_ops && !packed_ops)) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_cigar_indel") : SVB_ERR_ARG;
    // one mapped, forward, mapq-255 record at position 0 of a contig as long as a BAM coordinate can be
    std::vector<uint32_t> ops((static_cast<size_t>(n_ops) + 3) / 4 * 4, 15u);
    std::copy(packed_ops, packed_ops + n_ops, ops.begin());
    svb_aln_hdr h;
    memset(&h, 0, sizeof h);
    h.mapq = 255;
    h.n_cigar = n_ops;
    h.l_seq = 0xFFFFFFFFu;
    const int32_t clen = 0x7FFFFFFF, lex = 0;
    svb_records* rec = nullptr;
    int rc = svb_load_records(ctx, &h, 1, ops.data(), ops.size(), nullptr, nullptr, 0, &clen, &lex, 1, &rec);
    if (rc != SVB_OK) return rc;
    svb_params p;
    memset(&p, 0, sizeof p);
    p.min_sv_size = min_length;
    svb_table* t = nullptr;
    rc = svb_collect(ctx, rec, &p, 0, &t);
    svb_records_free(rec);
    if (rc != SVB_OK) return rc;
    std::vector<svb_row> rows(t->n);
    uint64_t n = 0;
    rc = svb_table_to_host(ctx, t, rows.data(), rows.size(), &n);
    svb_table_free(t);
    if (rc != SVB_OK) return rc;
    *n_out = static_cast<uint32_t>(n);
    for (uint64_t i = 0; i < n && i < cap; ++i) {
        const svb_row& r = rows[i];
        const bool del = r.type == SVB_DEL;
        // with pos = 0 and an unbounded contig the clamps are the identity: start == pos_ref
        out4[4 * i + 0] = del ? r.src_start : r.dst_start;
        out4[4 * i + 1] = static_cast<int64_t>(r.seq_pos);      // pos_read is kept for deletions too
        out4[4 * i + 2] = del ? static_cast<int64_t>(r.src_end) - r.src_start : static_cast<int64_t>(r.dst_end) - r.dst_start;
        out4[4 * i + 3] = del ? 1 : 0;
    }
    return SVB_OK;
}

// ---- reference genome ------------------------------------------------------------------------

int svb_ref_load(svb_ctx* ctx, const uint8_t* bases, const uint64_t* contig_off, int32_t n_contig, svb_ref** out) {
    if (!ctx || !out || n_contig < 0 || !contig_off) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_ref_load") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    svb_ref* r = new (std::nothrow) svb_ref();
    if (!r) return svb_fail(ctx, SVB_ERR_NOMEM, "svb_ref_load");
    r->device = ctx->device;
    r->n_contig = n_contig;
    r->n_bases = contig_off[n_contig];
    cudaError_t e = cudaMalloc(&r->d_contig_off, sizeof(uint64_t) * (static_cast<size_t>(n_contig) + 1));
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(r->d_contig_off, contig_off, sizeof(uint64_t) * (static_cast<size_t>(n_contig) + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && r->n_bases) e = cudaMalloc(&r->d_bases, r->n_bases);
    if (e == cudaSuccess && r->n_bases) e = cudaMemcpyAsync(r->d_bases, bases, r->n_bases, cudaMemcpyHostToDevice, ctx->stream);
    uint8_t map[256];
    std::string why;
    if (build_class_map(bases, r->n_bases, map, &why) != SVB_OK) {
        cudaStreamSynchronize(ctx->stream);
        svb_ref_free(r);
        return svb_fail(ctx, SVB_ERR_FORMAT, why.c_str());
    }
    if (e == cudaSuccess) e = cudaMalloc(&r->d_class_map, 256);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->d_class_map, map, 256, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        svb_ref_free(r);
        return svb_fail(ctx, SVB_ERR_CUDA, "svb_ref_load", e);
    }
    *out = r;
    return SVB_OK;
}

// test / debugging aid: the resident reference back on the host
int svb_ref_to_host(svb_ctx* ctx, const svb_ref* r, uint8_t* bases_dst, uint64_t cap, uint64_t* n_bases, uint8_t* class_map256_dst) {
    if (!ctx || !r || !n_bases) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_ref_to_host") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    *n_bases = r->n_bases;
    if (bases_dst) {
        if (cap < r->n_bases) return svb_fail(ctx, SVB_ERR_ARG, "svb_ref_to_host: destination too small");
        if (r->n_bases) SVB_CUDA(ctx, cudaMemcpyAsync(bases_dst, r->d_bases, r->n_bases, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (class_map256_dst) SVB_CUDA(ctx, cudaMemcpyAsync(class_map256_dst, r->d_class_map, 256, cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SVB_OK;
}

void svb_ref_free(svb_ref* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    cudaFree(r->d_bases);
    cudaFree(r->d_contig_off);
    cudaFree(r->d_class_map);
    delete r;
}

int svb_pair(svb_ctx* ctx, const svb_table* h1, const svb_table* h2, const svb_records* rec1, const svb_records* rec2,
             const svb_ref* ref, const svb_params* p, svb_table** out) {
    if (!ctx || !h1 || !h2 || !ref || !p || !out) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_pair") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    int rc = run_pairing(ctx, h1, h2, rec1, rec2, ref, p, out);
    if (rc != SVB_OK) return rc;
    if (*reinterpret_cast<const uint32_t*>(ctx->h_pinned + 12) == 0u) return SVB_OK;      // status word read back by run_pairing
    rc = check_device_status(ctx);
    if (rc != SVB_OK && *out) {
        svb_table_free(*out);
        *out = nullptr;
    }
    return rc;
}

int svb_edit_distance(svb_ctx* ctx, const uint8_t* a, const uint64_t* a_off, const uint8_t* b, const uint64_t* b_off,
                      uint32_t n_pairs, int64_t* out) {
    if (!ctx || !a_off || !b_off || (n_pairs && !out)) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_edit_distance") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    return run_edit_distance_strings(ctx, a, a_off, b, b_off, n_pairs, -1, out);
}

int svb_edit_distance_bounded(svb_ctx* ctx, const uint8_t* a, const uint64_t* a_off, const uint8_t* b, const uint64_t* b_off,
                              uint32_t n_pairs, int64_t max_distance, int64_t* out) {
    if (!ctx || !a_off || !b_off || (n_pairs && !out) || max_distance < 0)
        return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_edit_distance_bounded") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    return run_edit_distance_strings(ctx, a, a_off, b, b_off, n_pairs, max_distance, out);
}

int svb_form_partitions(svb_ctx* ctx, const uint64_t* keys, uint32_t n, int64_t max_distance, uint32_t* order,
                        uint32_t* part_start, uint32_t* n_parts) {
    if (!ctx || !n_parts || (n && (!keys || !order || !part_start))) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_form_partitions") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    return run_form_partitions(ctx, keys, n, max_distance, order, part_start, n_parts);
}

int svb_cluster_labels(svb_ctx* ctx, const double* condensed, const uint32_t* n_points, uint32_t n_problems,
                       double threshold, int32_t* labels_out) {
    if (!ctx || !n_points || !labels_out) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_cluster_labels") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    return run_cluster_labels(ctx, condensed, n_points, n_problems, threshold, labels_out);
}

}  // extern "C"
