// VCF body on the device: one warp per record line (vcf_core.cuh holds the per-record plan and the byte getter).
//
// Reference: write_final_vcf's record loop (SVIM_COMBINE.py:428-475) calling Candidate*.get_vcf_entry*; with sequence
// alleles every DEL / INS / INV line carries its REF and ALT bases, fetched one record at a time from the FASTA and the
// query sequence.  Here the bases are already in HBM (svb_ref, the record images' 4-bit sequences): the host only
// sizes the lines (no bases involved), the kernel gathers, and the finished text comes back in one copy.
#include <vector>

#include "common.cuh"
#include "vcf_core.cuh"

namespace {

constexpr int VCF_WARPS = 4;

__global__ void __launch_bounds__(VCF_WARPS * 32)
vcf_write_kernel(const svb_row* __restrict__ rows, const svb_vcf_entry* __restrict__ entries, const uint64_t* __restrict__ line_off,
                 uint32_t n_entries, VcfEnv env, uint8_t* __restrict__ out) {
    __shared__ VcfPlan plans[VCF_WARPS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t e = blockIdx.x * VCF_WARPS + warp;
    if (e >= n_entries) return;
    VcfPlan& p = plans[warp];
    if (lane == 0) {
        const svb_vcf_entry en = entries[e];
        vcf_plan(rows[en.row], en, env, p);
    }
    __syncwarp();
    uint8_t* dst = out + line_off[e];
    for (uint32_t k = 0; k < p.n_pieces; ++k) {
        const VcfPiece q = p.piece[k];
        for (uint64_t i = lane; i < q.len; i += 32ull) dst[i] = vcf_piece_byte(p, q, env, i);
        dst += q.len;
    }
}

bool mode_fits_type(uint32_t mode, uint8_t type) {
    switch (mode) {
        case VCF_DEL: return type == SVB_DEL;
        case VCF_INV: return type == SVB_INV;
        case VCF_INS: return type == SVB_INS;
        case VCF_TAN_AS_INS: case VCF_TAN_AS_DUP: return type == SVB_DUP_TAN;
        case VCF_INT_AS_INS: case VCF_INT_AS_DUP: return type == SVB_DUP_INT;
        case VCF_BND: case VCF_BND_MATE: return type == SVB_BND;
        default: return false;
    }
}

}  // namespace

extern "C" int svb_vcf_body(svb_ctx* ctx, const svb_table* t, const svb_records* const rec[3], const svb_ref* ref, const char* names,
                            const uint32_t* name_off, int32_t n_contig, const svb_vcf_entry* entries, uint64_t n_entries,
                            uint32_t flags, const uint8_t** text, uint64_t* n_bytes) {
    if (!ctx) return SVB_ERR_ARG;
    if (!t || !ref || !rec || !names || !name_off || !text || !n_bytes || (n_entries && !entries) || n_entries > 0xFFFFFFFFull)
        return svb_fail(ctx, SVB_ERR_ARG, "svb_vcf_body");
    if (n_contig != ref->n_contig) return svb_fail(ctx, SVB_ERR_ARG, "svb_vcf_body: the contig names do not match the reference");
    cudaSetDevice(ctx->device);
    *text = nullptr;
    *n_bytes = 0;
    if (n_entries == 0) return SVB_OK;

    // line lengths on the host: the plan of a record needs its row and the contig geometry, no bases
    std::vector<svb_row> rows(t->n);
    std::vector<uint64_t> contig_off(static_cast<size_t>(n_contig) + 1);
    if (t->n) SVB_CUDA(ctx, cudaMemcpyAsync(rows.data(), t->d_rows, sizeof(svb_row) * t->n, cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(contig_off.data(), ref->d_contig_off, sizeof(uint64_t) * contig_off.size(), cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    VcfEnv env = {};
    env.contig_off = contig_off.data();
    env.n_contig = n_contig;
    env.name_off = name_off;
    env.flags = flags;
    std::vector<uint64_t> line_off(n_entries + 1, 0);
    VcfPlan plan;
    for (uint64_t i = 0; i < n_entries; ++i) {
        const svb_vcf_entry& en = entries[i];
        if (en.row >= t->n || !mode_fits_type(en.mode, rows[en.row].type))
            return svb_fail(ctx, SVB_ERR_ARG, "svb_vcf_body: an entry names a row of another class (or no row)");
        const svb_row& r = rows[en.row];
        const int32_t tids[2] = {r.src_tid, r.dst_tid};
        const bool need[2] = {r.type != SVB_INS, r.type == SVB_INS || r.type == SVB_DUP_INT || r.type == SVB_BND};
        for (int k = 0; k < 2; ++k)
            if (need[k] && (tids[k] < 0 || tids[k] >= n_contig)) return svb_fail(ctx, SVB_ERR_ARG, "svb_vcf_body: contig id out of range");
        if (en.mode == VCF_INS && !(flags & VCF_SYMBOLIC) && r.seq_len) {
            const svb_records* q = r.hap < 3 ? rec[r.hap] : nullptr;
            if (!q || !q->d_seq4 || !q->d_seq_off || r.aln_idx >= q->n_aln)
                return svb_fail(ctx, SVB_ERR_ARG, "svb_vcf_body: an insertion needs query bases that are not resident (svb_records_set_sequences)");
        }
        vcf_plan(r, en, env, plan);
        line_off[i + 1] = line_off[i] + vcf_plan_length(plan);
    }
    const uint64_t total = line_off[n_entries];

    if (total > ctx->h_text_cap) {
        if (ctx->h_text) cudaFreeHost(ctx->h_text);
        ctx->h_text = nullptr;
        ctx->h_text_cap = 0;
        const size_t cap = static_cast<size_t>(total + total / 4 + 4096);
        SVB_CUDA(ctx, cudaMallocHost(&ctx->h_text, cap));
        ctx->h_text_cap = cap;
    }
    const uint32_t names_bytes = name_off[n_contig];
    struct Scratch {                     // stream-ordered device scratch, released on every way out
        cudaStream_t stream;
        std::vector<void*> ptrs;
        ~Scratch() {
            for (void* q : ptrs) cudaFreeAsync(q, stream);
        }
        cudaError_t get(void** out, size_t bytes) {
            const cudaError_t e = cudaMallocAsync(out, bytes, stream);
            if (e == cudaSuccess) ptrs.push_back(*out);
            return e;
        }
    } scratch{ctx->stream, {}};
    uint8_t* d_out = nullptr;
    uint8_t* d_names = nullptr;
    uint32_t* d_name_off = nullptr;
    svb_vcf_entry* d_entries = nullptr;
    uint64_t* d_line_off = nullptr;
    SVB_CUDA(ctx, scratch.get(reinterpret_cast<void**>(&d_out), std::max<uint64_t>(total, 1)));
    SVB_CUDA(ctx, scratch.get(reinterpret_cast<void**>(&d_names), std::max<uint32_t>(names_bytes, 1u)));
    SVB_CUDA(ctx, scratch.get(reinterpret_cast<void**>(&d_name_off), sizeof(uint32_t) * (static_cast<size_t>(n_contig) + 1)));
    SVB_CUDA(ctx, scratch.get(reinterpret_cast<void**>(&d_entries), sizeof(svb_vcf_entry) * n_entries));
    SVB_CUDA(ctx, scratch.get(reinterpret_cast<void**>(&d_line_off), sizeof(uint64_t) * (n_entries + 1)));
    if (names_bytes) SVB_CUDA(ctx, cudaMemcpyAsync(d_names, names, names_bytes, cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(d_name_off, name_off, sizeof(uint32_t) * (static_cast<size_t>(n_contig) + 1), cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(d_entries, entries, sizeof(svb_vcf_entry) * n_entries, cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(d_line_off, line_off.data(), sizeof(uint64_t) * (n_entries + 1), cudaMemcpyHostToDevice, ctx->stream));
    VcfEnv dev = {};
    dev.bases = ref->d_bases;
    dev.contig_off = ref->d_contig_off;
    dev.n_contig = n_contig;
    dev.names = d_names;
    dev.name_off = d_name_off;
    for (int h = 0; h < 3; ++h) {
        dev.seq4[h] = rec[h] ? rec[h]->d_seq4 : nullptr;
        dev.seq_off[h] = rec[h] ? rec[h]->d_seq_off : nullptr;
    }
    dev.flags = flags;
    {
        KernelTimer timer(ctx, SVB_K_VCF);
        const unsigned blocks = static_cast<unsigned>((n_entries + VCF_WARPS - 1) / VCF_WARPS);
        vcf_write_kernel<<<blocks, VCF_WARPS * 32, 0, ctx->stream>>>(t->d_rows, d_entries, d_line_off, static_cast<uint32_t>(n_entries), dev, d_out);
        ctx->launches += 1;
    }
    SVB_CUDA(ctx, cudaGetLastError());
    if (total) SVB_CUDA(ctx, cudaMemcpyAsync(ctx->h_text, d_out, total, cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *text = ctx->h_text;
    *n_bytes = total;
    return SVB_OK;
}
