// Reference genome straight from the FASTA file to HBM (the part of pysam.FastaFile the pairing stage needs:
// svim-asm:124, SVIM_COMBINE.py:45-99 reference.fetch).  The file is mapped and copied to the device as it is; one
// kernel drops the line terminators (faidx geometry: offset, linebases, linewidth per contig), upper-cases
// (compute_distance calls .upper(), SVIM_COMBINE.py:47-99) and records which byte values occur (symbol classes of the
// edit-distance kernel).  The host never touches the 3 GB of bases.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "pairing.cuh"

namespace {

struct FaiEntry {               // one contig, in the order of the BAM header
    uint64_t length, offset, linebases, linewidth;
    uint64_t out_off;           // first base in the concatenated array
};

// grid-stride over the output bases of one contig (blockIdx.y = contig)
__global__ void fasta_compact_kernel(const uint8_t* __restrict__ raw, uint64_t raw_size, const FaiEntry* __restrict__ fai,
                                     uint8_t* __restrict__ bases, uint32_t* __restrict__ seen_bits, uint32_t* status) {
    const FaiEntry c = fai[blockIdx.y];
    uint32_t seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < c.length; i += stride) {
        const uint64_t src = c.offset + (i / c.linebases) * c.linewidth + i % c.linebases;
        uint8_t b = 0;
        if (src < raw_size) b = raw[src];
        else atomicOr(status, 1u);                           // the index points past the end of the file
        if (b >= 'a' && b <= 'z') b -= 32;
        bases[c.out_off + i] = b;
        seen[b >> 5] |= 1u << (b & 31u);
    }
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t v = __reduce_or_sync(0xffffffffu, seen[w]);
        if ((threadIdx.x & 31u) == 0 && v) atomicOr(seen_bits + w, v);
    }
}

}  // namespace

extern "C" {

// fai: n_contig rows of {length, offset, linebases, linewidth} in BAM header order; length 0 = contig absent from the FASTA
int svb_ref_load_fasta(svb_ctx* ctx, const char* path, const uint64_t* fai, int32_t n_contig, svb_ref** out) {
    if (!ctx || !path || !out || n_contig < 0 || (n_contig && !fai)) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_ref_load_fasta") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return svb_fail(ctx, SVB_ERR_IO, "svb_ref_load_fasta: cannot open the FASTA file");
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size <= 0) {
        close(fd);
        return svb_fail(ctx, SVB_ERR_IO, "svb_ref_load_fasta: cannot stat the FASTA file");
    }
    const size_t fsize = static_cast<size_t>(sb.st_size);
    std::vector<FaiEntry> entries(static_cast<size_t>(n_contig));
    std::vector<uint64_t> contig_off(static_cast<size_t>(n_contig) + 1, 0);
    for (int32_t i = 0; i < n_contig; ++i) {
        FaiEntry& e = entries[static_cast<size_t>(i)];
        e.length = fai[4 * i];
        e.offset = fai[4 * i + 1];
        e.linebases = fai[4 * i + 2] ? fai[4 * i + 2] : 1;
        e.linewidth = fai[4 * i + 3] ? fai[4 * i + 3] : 1;
        e.out_off = contig_off[static_cast<size_t>(i)];
        contig_off[static_cast<size_t>(i) + 1] = e.out_off + e.length;
    }
    svb_ref* r = new (std::nothrow) svb_ref();
    if (!r) {
        close(fd);
        return svb_fail(ctx, SVB_ERR_NOMEM, "svb_ref_load_fasta");
    }
    r->device = ctx->device;
    r->n_contig = n_contig;
    r->n_bases = contig_off[static_cast<size_t>(n_contig)];
    uint8_t* d_raw = nullptr;
    FaiEntry* d_fai = nullptr;
    uint32_t* d_seen = nullptr;      // [8] presence bits, [8] status
    uint32_t h_seen[9] = {0};
    // Every allocation is stream-ordered (cudaMallocAsync / cudaFreeAsync): this load runs in the background of the BAM ingests,
    // and a plain cudaFree waits for the whole device -- for the other thread's inflate kernel -- holding up that thread's own
    // allocations meanwhile (seen as stalls of 50-700 ms in the ingest's allocation phase).
    cudaStream_t st = ctx->stream;
    cudaError_t e = cudaMallocAsync(&d_raw, fsize, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);                   // (the upload's streams are not ordered behind `st`)
    if (e == cudaSuccess && upload_file_range(ctx, fd, 0, fsize, d_raw, true, true) != SVB_OK) {      // gives way to a BAM upload in flight: the ingest is waiting for that one
        close(fd);
        cudaFreeAsync(d_raw, st);
        svb_ref_free(r);
        return SVB_ERR_IO;
    }
    if (e == cudaSuccess) e = cudaMallocAsync(&d_fai, sizeof(FaiEntry) * std::max<size_t>(entries.size(), 1), st);
    if (e == cudaSuccess && n_contig) e = cudaMemcpyAsync(d_fai, entries.data(), sizeof(FaiEntry) * entries.size(), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_seen, sizeof(uint32_t) * 9, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_seen, 0, sizeof(uint32_t) * 9, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&r->d_contig_off, sizeof(uint64_t) * contig_off.size(), st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->d_contig_off, contig_off.data(), sizeof(uint64_t) * contig_off.size(), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && r->n_bases) e = cudaMallocAsync(&r->d_bases, r->n_bases, st);
    if (e == cudaSuccess && r->n_bases && n_contig) {
        const dim3 grid(static_cast<unsigned>(ctx->sm_count) * 4u, static_cast<unsigned>(n_contig));
        fasta_compact_kernel<<<grid, 256, 0, st>>>(d_raw, fsize, d_fai, r->d_bases, d_seen, d_seen + 8);
        ctx->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_seen, d_seen, sizeof h_seen, cudaMemcpyDeviceToHost, st);
    if (d_raw) cudaFreeAsync(d_raw, st);
    if (d_fai) cudaFreeAsync(d_fai, st);
    if (d_seen) cudaFreeAsync(d_seen, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    close(fd);
    if (e != cudaSuccess) {
        svb_ref_free(r);
        return svb_fail(ctx, SVB_ERR_CUDA, "svb_ref_load_fasta", e);
    }
    if (h_seen[8]) {
        svb_ref_free(r);
        return svb_fail(ctx, SVB_ERR_FORMAT, "svb_ref_load_fasta: the index points past the end of the FASTA file");
    }
    bool seen[256];
    for (int c = 0; c < 256; ++c) seen[c] = (h_seen[c >> 5] >> (c & 31)) & 1u;
    uint8_t map[256];
    std::string why;
    if (build_class_map_from_seen(seen, map, &why) != SVB_OK) {
        svb_ref_free(r);
        return svb_fail(ctx, SVB_ERR_FORMAT, why.c_str());
    }
    e = cudaMallocAsync(&r->d_class_map, 256, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->d_class_map, map, 256, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        svb_ref_free(r);
        return svb_fail(ctx, SVB_ERR_CUDA, "svb_ref_load_fasta: class map", e);
    }
    *out = r;
    return SVB_OK;
}

}  // extern "C"
