// Device ingest: BGZF/BAM file -> the record image of svb_load_records, built ON the GPU.
//
// Stands in for what the reference obtains from pysam/htslib on the host (svim-asm:63,85-86 AlignmentFile;
// SVIM_COLLECT.py:65 bam.fetch; SVIM_intra.py:37 cigartuples incl. the CG:B,I long-CIGAR convention).  The host only
// reads the file, walks the BGZF member headers (18 bytes each) and parses the few hundred KB of names and SA tags;
// the compressed bytes cross PCIe once and everything else happens in HBM:
//   bgzf_inflate_kernel   one warp per BGZF member (independent raw-deflate streams; decoder of inflate_core.cuh in lane 0,
//                         matches and stored blocks copied by the whole warp)
//   bam_chase_kernel      record boundaries (each record starts where the previous one ends: one dependent load per record)
//   bam_fields_kernel     one thread per record: fixed fields, tag walk (SA:Z, CG:B,I), source offsets
//   bam_copy_kernel       one CTA per record: CIGAR ops into 16-byte aligned runs padded with op 15, 4-bit query
//                         bases, read names and SA texts into flat buffers
// The result is bit-identical to the host ingest (bam_ingest.cpp): same svb_aln_hdr array, same CIGAR / sequence
// layout, same SA segments (the SA text itself is parsed by the same host routine, svb_parse_sa).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "bam_host.h"
#include "common.cuh"
#include "inflate_core.cuh"

extern "C" int load_records_impl(svb_ctx* ctx, const svb_aln_hdr* hdr, uint32_t n_aln, const uint32_t* cigar, uint4* d_cigar_prebuilt,
                                 uint64_t n_ops_padded, const svb_segment* seg, const uint32_t* sa_count, uint32_t n_seg,
                                 const int32_t* contig_len, const int32_t* contig_lexrank, int32_t n_contig, svb_records** out);

namespace {

enum : uint32_t {
    ING_ERR_INFLATE = 1u,       // a member did not inflate to its ISIZE
    ING_ERR_RECORD = 2u,        // truncated / malformed BAM record
    ING_ERR_TAG = 4u,           // malformed auxiliary field
    ING_ERR_INDEX = 8u          // the .bai offsets do not lie on the record chain (stale index): serial walk instead
};

struct DevMember {
    uint64_t in_off, out_off;
    uint32_t in_len, out_len;
};

// One WARP per member.  A thread per member would put 32 unrelated decoder states into one warp: they diverge at every
// branch and the warp runs them one after the other (measured: 2.5 GB/s).  A warp whose lane 0 decodes while the others
// wait for matches to copy executes 110 instructions per symbol, most of them the decoding lane's byte loops (35 GB/s).
// Here EVERY lane decodes the SAME member in lock step (inflate_core.cuh: inf_run_lanes) -- the same instruction stream
// costs a warp the same whether one lane or all 32 execute it -- so every lane knows every symbol without a broadcast and
// the bytes of a match are copied by the lanes side by side.  Lane 0 alone parses the block headers (they build the code
// arrays in shared memory) and stores the literals.  Members of one CTA never interact.
constexpr int INF_WARPS = 4;
#ifndef INF_MIN_CTAS
#define INF_MIN_CTAS 9          // resident CTAs per SM the kernel is compiled for (registers <= 65536 / (INF_MIN_CTAS * 128))
#endif
// 5.8 KB per warp: 9 CTAs of 4 warps fit the SM's 227 KB.  The scratch array of the block header (code lengths, 320 bytes) lives
// in the distance table: the tables are built after the header has been parsed.
struct InfTables {
    uint16_t lencnt[16], lensym[288], distcnt[16], distsym[32];
    // direct lookup on the next bits of the stream (entries of inflate_core.cuh: literal, or base + extra-bit count)
    inf_len_t fast_len[1 << INF_LEN_BITS];
    uint32_t fast_dist[1 << INF_DIST_BITS];
};
static_assert(sizeof(uint32_t) * (1 << INF_DIST_BITS) >= 320, "the header's scratch array must fit the distance table");

// One warp, one member: block headers by lane 0 (the reader's state is handed to every lane afterwards), the direct tables
// filled by all lanes, then the symbol loop in lock step; stored blocks go through the ring in pieces.
__device__ int inflate_member_warp(const uint8_t* __restrict__ src, uint32_t src_len, uint8_t* dst, uint32_t out_len, InfTables& T,
                                   uint8_t* ring, uint32_t lane) {
    constexpr uint32_t FULLM = 0xffffffffu;
    InfHuff lencode{T.lencnt, T.lensym}, distcode{T.distcnt, T.distsym};
    InfBits b{src, src + src_len, 0ull, 0, 0};          // the same on every lane
    InfOut o = inf_out(ring, dst, out_len);             // the same on every lane
    int last = 0;
    do {
        int err = 0;
        uint32_t type = 0, st_off = 0, st_len = 0;
        if (lane == 0) err = inf_block_header(b, src, o.pos, out_len, lencode, distcode, reinterpret_cast<uint8_t*>(T.fast_dist), &last, &type, &st_off, &st_len);
        err = __shfl_sync(FULLM, err, 0);
        if (err) return err;
        type = __shfl_sync(FULLM, type, 0);
        last = __shfl_sync(FULLM, last, 0);
        {                                               // lane 0's reader, to every lane
            const uint32_t at = __shfl_sync(FULLM, static_cast<uint32_t>(b.p - src), 0);
            const uint32_t lo = __shfl_sync(FULLM, static_cast<uint32_t>(b.buf), 0), hi = __shfl_sync(FULLM, static_cast<uint32_t>(b.buf >> 32), 0);
            b.p = src + at;
            b.buf = (static_cast<uint64_t>(hi) << 32) | lo;
            b.cnt = __shfl_sync(FULLM, b.cnt, 0);
            b.virt = __shfl_sync(FULLM, b.virt, 0);
        }
        if (type == 0u) {                               // stored block: through the ring in pieces (later matches may point into it)
            st_off = __shfl_sync(FULLM, st_off, 0);
            st_len = __shfl_sync(FULLM, st_len, 0);
            for (uint32_t done = 0; done < st_len;) {
                const uint32_t chunk = min(st_len - done, INF_FLUSH_AT);
                for (uint32_t i = lane; i < chunk; i += 32u) ring[(o.rbase + o.pos + i) & INF_RMASK] = src[st_off + done + i];
                o.pos += chunk;
                done += chunk;
                __syncwarp();
                inf_flush(o, o.pos, lane, 32u);
                o.flushed = o.pos;
                __syncwarp();
            }
            continue;
        }
        for (uint32_t i = lane; i < (1u << INF_LEN_BITS); i += 32u) T.fast_len[i] = 0u;
        for (uint32_t i = lane; i < (1u << INF_DIST_BITS); i += 32u) T.fast_dist[i] = 0u;
        __syncwarp();                                   // (with the shuffles above: lane 0's code arrays are visible)
        inf_fill_table(T.lencnt, T.lensym, T.fast_len, INF_LEN_BITS, false, lane, 32u);
        inf_fill_table(T.distcnt, T.distsym, T.fast_dist, INF_DIST_BITS, true, lane, 32u);
        __syncwarp();
        err = inf_run_lanes(b, lencode, distcode, T.fast_len, T.fast_dist, o, lane, 32u);
        if (err) return err;                            // the same on every lane
        __syncwarp();                                   // the block's last stores, before lane 0 rebuilds the tables
    } while (!last);
    if (o.pos > out_len) return INF_ERR_OUTPUT;
    __syncwarp();
    inf_flush(o, o.pos, lane, 32u);
    return o.pos == out_len ? INF_OK : INF_ERR_SIZE;
}

// Persistent warps: every warp claims the next member from a counter until none is left, so a slow member never holds
// the other three warps of its CTA (and the SM's slot) idle.  status[2..3]: sum of member cycles, status[4]: the counter.
// (A variant that was launched before the upload and waited, member by member, for "chunk landed" flags written by the copy
// streams overlapped the two -- and hung for minutes when another host thread made a device-synchronising call while the
// kernel was spinning: a kernel must not depend on host progress.  The upload now overlaps the host's member table instead.)
__global__ void __launch_bounds__(INF_WARPS * 32, INF_MIN_CTAS) bgzf_inflate_kernel(const uint8_t* __restrict__ comp, const DevMember* __restrict__ members,
                                                                                    uint32_t n, uint8_t* out, uint32_t* status) {
    __shared__ InfTables tables[INF_WARPS];
    __shared__ __align__(16) uint8_t rings[INF_WARPS][INF_RING];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    while (true) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(status + 4, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) break;
        const DevMember m = members[i];
        const long long t0 = clock64();
        const int rc = inflate_member_warp(comp + m.in_off, m.in_len, out + m.out_off, m.out_len, tables[warp], rings[warp], lane);
        if (lane == 0) {
            if (rc != INF_OK) atomicOr(status, ING_ERR_INFLATE);
            atomicAdd(reinterpret_cast<unsigned long long*>(status + 2), static_cast<unsigned long long>(clock64() - t0));
        }
        __syncwarp();
    }
}

__device__ __forceinline__ uint32_t ld_u32(const uint8_t* p) {        // unaligned little-endian loads
    return p[0] | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) | (static_cast<uint32_t>(p[3]) << 24);
}
__device__ __forceinline__ uint32_t ld_u16(const uint8_t* p) { return p[0] | (static_cast<uint32_t>(p[1]) << 8); }

// body offset of every record; every hop is one dependent load (block_size), at least 36 bytes forward
__global__ void bam_chase_kernel(const uint8_t* __restrict__ data, uint64_t start, uint64_t end, uint64_t* __restrict__ rec_off,
                                 uint32_t cap, uint32_t* n_rec, uint32_t* status) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint64_t p = start;
    uint32_t n = 0;
    while (p + 4 <= end) {
        const int32_t block = static_cast<int32_t>(ld_u32(data + p));
        if (block < 32 || p + 4 + static_cast<uint64_t>(block) > end) {
            atomicOr(status, ING_ERR_RECORD);
            break;
        }
        if (n < cap) rec_off[n] = p + 4;
        ++n;
        p += 4 + static_cast<uint64_t>(block);
    }
    *n_rec = n;
}

// The same walk in parallel.  A BAM file comes with an index (the reference insists on it: svim-asm:66-72 check_index), and the
// index names record starts: every chunk begin of every bin and every linear-index entry is the virtual offset of a record.
// For genome-genome alignments (records of hundreds of kilobases) that is nearly every record.  Thread i walks from known
// start i to known start i + 1 and MUST land on it exactly; anything else (a stale or foreign index) raises ING_ERR_INDEX and
// the host falls back to the serial walk.  WRITE = false counts the records of every segment, WRITE = true writes their
// offsets at the scanned positions.
template <bool WRITE>
__global__ void bam_chase_indexed_kernel(const uint8_t* __restrict__ data, const uint64_t* __restrict__ starts, uint32_t n_seg,
                                         uint32_t* __restrict__ counts, uint64_t* __restrict__ rec_off, uint32_t* status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_seg) return;
    uint64_t p = starts[i];
    const uint64_t end = starts[i + 1];
    uint32_t n = 0;
    const uint32_t base = WRITE ? counts[i] : 0u;
    while (p < end) {
        if (p + 4 > end) { atomicOr(status, ING_ERR_INDEX); break; }
        const int32_t block = static_cast<int32_t>(ld_u32(data + p));
        if (block < 32 || p + 4 + static_cast<uint64_t>(block) > end) { atomicOr(status, ING_ERR_INDEX); break; }
        if (WRITE) rec_off[base + n] = p + 4;
        ++n;
        p += 4 + static_cast<uint64_t>(block);
    }
    if (!WRITE) counts[i] = n;
}

struct RecInfo {                 // where the variable-length parts of one record sit in the inflated stream
    uint64_t cigar_src;          // first real op (inside the record, or the CG:B,I payload)
    uint64_t seq_src;
    uint64_t sa_src;             // SA:Z value (without the NUL), sa_len == 0xFFFFFFFF if absent
    uint64_t name_src;
    uint32_t n_cigar, l_seq, sa_len, name_len;
};

__global__ void bam_fields_kernel(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off, uint32_t n,
                                  svb_aln_hdr* __restrict__ hdr, RecInfo* __restrict__ info, uint32_t* status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t off = rec_off[i];
    const uint8_t* body = data + off;
    const uint32_t block = ld_u32(body - 4);
    const uint32_t l_name = body[8];
    const uint32_t n_cig = ld_u16(body + 12);
    const int32_t l_seq = static_cast<int32_t>(ld_u32(body + 16));
    RecInfo r;
    r.name_src = off + 32;
    r.name_len = l_name;
    const uint64_t cig = off + 32 + l_name;
    r.seq_src = cig + 4ull * n_cig;
    const uint64_t tags = r.seq_src + (static_cast<uint64_t>(l_seq < 0 ? 0 : l_seq) + 1) / 2 + static_cast<uint64_t>(l_seq < 0 ? 0 : l_seq);
    const uint64_t rec_end = off + block;
    r.sa_src = 0;
    r.sa_len = 0xFFFFFFFFu;
    uint64_t cg_src = 0;
    uint32_t cg_n = 0;
    bool have_cg = false;
    if (l_seq < 0 || tags > rec_end) {
        atomicOr(status, ING_ERR_RECORD);
    } else {
        uint64_t t = tags;
        while (t + 3 <= rec_end) {               // every round advances by at least 4 bytes
            const uint8_t a0 = data[t], a1 = data[t + 1], ty = data[t + 2];
            t += 3;
            uint64_t adv = 0;
            bool bad = false;
            switch (ty) {
                case 'A': case 'c': case 'C': adv = 1; break;
                case 's': case 'S': adv = 2; break;
                case 'i': case 'I': case 'f': adv = 4; break;
                case 'Z': case 'H': {
                    uint64_t z = t;
                    while (z < rec_end && data[z] != 0) ++z;
                    if (z >= rec_end) { bad = true; break; }
                    if (a0 == 'S' && a1 == 'A' && ty == 'Z') {
                        r.sa_src = t;
                        r.sa_len = static_cast<uint32_t>(z - t);
                    }
                    adv = z - t + 1;
                    break;
                }
                case 'B': {
                    if (t + 5 > rec_end) { bad = true; break; }
                    const uint8_t sub = data[t];
                    const uint32_t cnt = ld_u32(data + t + 1);
                    const uint64_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                    if (a0 == 'C' && a1 == 'G' && sub == 'I') {
                        cg_src = t + 5;
                        cg_n = cnt;
                        have_cg = true;
                    }
                    adv = 5 + es * cnt;
                    break;
                }
                default: bad = true; break;
            }
            if (bad || t + adv > rec_end) {
                atomicOr(status, ING_ERR_TAG);
                break;
            }
            t += adv;
        }
    }
    // long CIGAR convention: "<l_seq>S<ref_len>N" placeholder + CG:B,I (htslib restores the real CIGAR transparently)
    r.cigar_src = cig;
    r.n_cigar = n_cig;
    if (have_cg && n_cig == 2u && ld_u32(data + cig) == ((static_cast<uint32_t>(l_seq) << 4) | 4u) && (ld_u32(data + cig + 4) & 15u) == 3u) {
        r.cigar_src = cg_src;
        r.n_cigar = cg_n;
    }
    r.l_seq = static_cast<uint32_t>(l_seq);
    info[i] = r;
    svb_aln_hdr h;
    h.tid = static_cast<int32_t>(ld_u32(body));
    h.pos = static_cast<int32_t>(ld_u32(body + 4));
    h.flag = static_cast<uint16_t>(ld_u16(body + 14));
    h.mapq = body[9];
    h.reserved0 = 0;
    h.n_cigar = r.n_cigar;
    h.cigar_off = 0;             // filled in on the host (prefix sums)
    h.l_seq = r.l_seq;
    h.sa_first = 0;
    hdr[i] = h;
}

struct CopyDst {                 // destination offsets of one record (host prefix sums)
    uint64_t cigar_op;           // index of its first op in cigar[] (multiple of 4)
    uint64_t seq_byte, name_byte, sa_byte;
};

// one CTA per record: the three variable-length parts move to their flat arrays
__global__ void __launch_bounds__(256) bam_copy_kernel(const uint8_t* __restrict__ data, const RecInfo* __restrict__ info,
                                                       const CopyDst* __restrict__ dst, uint32_t n, uint32_t* __restrict__ cigar,
                                                       uint8_t* __restrict__ seq4, char* __restrict__ names, char* __restrict__ sa_text) {
    const uint32_t i = blockIdx.x;
    if (i >= n) return;
    const RecInfo r = info[i];
    const CopyDst d = dst[i];
    const uint32_t padded = (r.n_cigar + 3u) / 4u * 4u;
    for (uint32_t k = threadIdx.x; k < padded; k += blockDim.x)
        cigar[d.cigar_op + k] = k < r.n_cigar ? ld_u32(data + r.cigar_src + 4ull * k) : 15u;
    const uint32_t seq_bytes = (r.l_seq + 1u) / 2u;
    if (seq4)
        for (uint32_t k = threadIdx.x; k < seq_bytes; k += blockDim.x) seq4[d.seq_byte + k] = data[r.seq_src + k];
    for (uint32_t k = threadIdx.x; k < r.name_len; k += blockDim.x) names[d.name_byte + k] = static_cast<char>(data[r.name_src + k]);
    if (r.sa_len != 0xFFFFFFFFu)
        for (uint32_t k = threadIdx.x; k <= r.sa_len; k += blockDim.x)
            sa_text[d.sa_byte + k] = k < r.sa_len ? static_cast<char>(data[r.sa_src + k]) : '\0';
}

void set_err(char* err, int err_len, const std::string& msg) {
    if (err && err_len > 0) snprintf(err, static_cast<size_t>(err_len), "%s", msg.c_str());
}

struct Cleanup {                 // frees whatever was allocated when the function leaves early
    cudaStream_t stream;
    std::vector<void*> dev;
    ~Cleanup() {
        for (void* p : dev)
            if (p) cudaFreeAsync(p, stream);
    }
};

}  // namespace

extern "C" {

// Record starts named by <path>.bai as offsets into the inflated stream, sorted, unique, inside (first_record, total_out).
// Empty when there is no usable index (the caller then walks the chain serially).
static std::vector<uint64_t> bai_record_starts(const std::string& bai_path, const std::vector<BgzfMember>& members, uint64_t first_record,
                                               uint64_t total_out) {
    std::vector<uint64_t> out;
    FILE* f = fopen(bai_path.c_str(), "rb");
    if (!f) return out;
    std::vector<uint8_t> buf;
    {
        fseek(f, 0, SEEK_END);
        const long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (n > 8 && n < (1l << 30)) {
            buf.resize(static_cast<size_t>(n));
            if (fread(buf.data(), 1, buf.size(), f) != buf.size()) buf.clear();
        }
        fclose(f);
    }
    if (buf.size() < 8 || memcmp(buf.data(), "BAI\1", 4) != 0) return out;
    const uint8_t* p = buf.data() + 4;
    const uint8_t* const end = buf.data() + buf.size();
    auto rd32 = [&](uint32_t* v) { if (end - p < 4) return false; memcpy(v, p, 4); p += 4; return true; };
    auto rd64 = [&](uint64_t* v) { if (end - p < 8) return false; memcpy(v, p, 8); p += 8; return true; };
    std::vector<uint64_t> voff;
    uint32_t n_ref = 0;
    if (!rd32(&n_ref)) return out;
    for (uint32_t r = 0; r < n_ref; ++r) {
        uint32_t n_bin = 0;
        if (!rd32(&n_bin)) return out;
        for (uint32_t b = 0; b < n_bin; ++b) {
            uint32_t bin = 0, n_chunk = 0;
            if (!rd32(&bin) || !rd32(&n_chunk)) return out;
            for (uint32_t c = 0; c < n_chunk; ++c) {
                uint64_t beg = 0, fin = 0;
                if (!rd64(&beg) || !rd64(&fin)) return out;
                if (bin != 37450u) voff.push_back(beg);          // 37450: the metadata pseudo-bin, not offsets of records
            }
        }
        uint32_t n_intv = 0;
        if (!rd32(&n_intv)) return out;
        for (uint32_t k = 0; k < n_intv; ++k) {
            uint64_t io = 0;
            if (!rd64(&io)) return out;
            if (io) voff.push_back(io);
        }
    }
    std::sort(voff.begin(), voff.end());
    voff.erase(std::unique(voff.begin(), voff.end()), voff.end());
    size_t m = 0;                                                // members are in file order: one forward scan serves all offsets
    for (uint64_t v : voff) {
        const uint64_t coff = v >> 16, uoff = v & 0xFFFFull;
        while (m < members.size() && members[m].file_off < coff) ++m;
        if (m == members.size()) break;
        if (members[m].file_off != coff || uoff >= members[m].out_len) continue;
        const uint64_t at = members[m].out_off + uoff;
        if (at > first_record && at < total_out) out.push_back(at);
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return out;
}

// timings of the last svb_bam_open_device call (ms): read file, H2D, inflate, chase, fields + copy, host parse, total
// (kept per context: two contexts ingesting at the same time do not share a buffer)
int svb_bam_device_timings(svb_ctx* ctx, double out[12]) {
    if (!ctx || !out) return SVB_ERR_ARG;
    for (int i = 0; i < 12; ++i) out[i] = ctx->ingest_ms[i];
    return SVB_OK;
}

int svb_bam_open_device(svb_ctx* ctx, const char* path, int keep_sequences, const int32_t* contig_lexrank_or_null, svb_bam** bam_out,
                        svb_records** rec_out, char* err, int err_len) {
    if (!ctx || !path || !bam_out || !rec_out) return SVB_ERR_ARG;
    *bam_out = nullptr;
    *rec_out = nullptr;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    Cleanup gc;
    gc.stream = st;
    cudaEvent_t ev[6];
    for (auto& e : ev) cudaEventCreate(&e);
    auto fail = [&](int code, const std::string& msg) {
        cudaStreamSynchronize(st);
        for (auto& e : ev) cudaEventDestroy(e);
        set_err(err, err_len, msg);
        return svb_fail(ctx, code, msg.c_str());
    };
    const auto wall0 = std::chrono::steady_clock::now();
    // SVB_INGEST_TRACE=1: host wall clock at every phase boundary, to stderr (where does the time between the kernels go?)
    const bool trace = getenv("SVB_INGEST_TRACE") != nullptr;
    auto mark = [&](const char* what) {
        if (trace) {
            const char* base = strrchr(path, '/');
            fprintf(stderr, "[ingest %s] %8.2f ms  %s\n", base ? base + 1 : path, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count(), what);
        }
    };

    // ---- the file: mapped, not read.  The only full pass over its bytes is the driver's copy to the device (the member
    // table touches 18 bytes per 64 KB); pinning gigabytes first would cost more than it saves.
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(SVB_ERR_IO, std::string("cannot open ") + path);
    struct stat sb;
    if (fstat(fd, &sb) != 0) {
        close(fd);
        return fail(SVB_ERR_IO, std::string("cannot stat ") + path);
    }
    const long fsize = static_cast<long>(sb.st_size);
    struct Mapping {
        void* p = MAP_FAILED;
        size_t n = 0;
        int fd = -1;
        ~Mapping() {
            if (p != MAP_FAILED) munmap(p, n);
            if (fd >= 0) close(fd);
        }
    } mapping;
    mapping.fd = fd;
    if (fsize <= 0) return fail(SVB_ERR_IO, "not a BGZF file");
    mapping.n = static_cast<size_t>(fsize);
    mapping.p = mmap(nullptr, mapping.n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (mapping.p == MAP_FAILED) return fail(SVB_ERR_IO, std::string("cannot map ") + path);
    madvise(mapping.p, mapping.n, MADV_SEQUENTIAL);
    const uint8_t* raw = static_cast<const uint8_t*>(mapping.p);
    const auto wall_read = std::chrono::steady_clock::now();

#define ING_CUDA(call)                                                             \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) return fail(SVB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)
    // ---- the upload starts at once, on a thread of its own (upload_file_range: parallel pread + pinned staging); it needs the
    // file's size only.  The host builds the member table meanwhile.
    uint8_t *d_comp = nullptr, *d_data = nullptr;
    DevMember* d_members = nullptr;
    uint32_t* d_status = nullptr;      // [0] error bits, [1] record count
    ING_CUDA(cudaMallocAsync(&d_comp, static_cast<size_t>(fsize), st));
    gc.dev.push_back(d_comp);
    ING_CUDA(cudaStreamSynchronize(st));                                   // (the upload's streams are not ordered behind `st`)
    struct Upload {
        std::thread worker;
        int rc = SVB_OK;
        double ms = 0.0;
        ~Upload() {
            if (worker.joinable()) worker.join();
        }
    } upload;
    upload.worker = std::thread([&upload, ctx, fd, fsize, d_comp]() {
        cudaSetDevice(ctx->device);
        const auto w0 = std::chrono::steady_clock::now();
        upload.rc = upload_file_range(ctx, fd, 0, static_cast<uint64_t>(fsize), d_comp, false, false);
        upload.ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    });

    // ---- meanwhile: the member table (18 bytes of every member's header through the mapping: 20 ms of page faults for 43,000 members)
    std::vector<BgzfMember> members;
    uint64_t total_out = 0;
    unsigned long long inflate_cycles = 0;
    {
        std::string why;
        const unsigned hw = std::thread::hardware_concurrency();
        if (!bgzf_member_table(raw, static_cast<uint64_t>(fsize), &members, &total_out, &why, getenv("SVB_INGEST_SERIAL_TABLE") ? 1 : static_cast<int>(std::min(8u, std::max(1u, hw / 4u)))))
            return fail(SVB_ERR_IO, why);
    }
    if (members.empty() || members.size() > 0x7fffffffull) return fail(SVB_ERR_IO, "not a BAM file");
    mark("member table");
    std::vector<DevMember> dm(members.size());
    for (size_t i = 0; i < members.size(); ++i) {
        dm[i].in_off = members[i].in_off;
        dm[i].out_off = members[i].out_off;
        dm[i].in_len = static_cast<uint32_t>(members[i].in_len);
        dm[i].out_len = static_cast<uint32_t>(members[i].out_len);
    }
    ING_CUDA(cudaMallocAsync(&d_data, std::max<uint64_t>(total_out, 1), st));
    gc.dev.push_back(d_data);
    ING_CUDA(cudaMallocAsync(&d_members, sizeof(DevMember) * dm.size(), st));
    gc.dev.push_back(d_members);
    ING_CUDA(cudaMallocAsync(&d_status, 8 * sizeof(uint32_t), st));      // [2..3]: 64-bit sum of member cycles, [4]: member counter
    gc.dev.push_back(d_status);
    ING_CUDA(cudaMemsetAsync(d_status, 0, 8 * sizeof(uint32_t), st));
    ING_CUDA(cudaMemcpyAsync(d_members, dm.data(), sizeof(DevMember) * dm.size(), cudaMemcpyHostToDevice, st));
    cudaEventRecord(ev[0], st);
    mark("allocations");
    const uint32_t n_members = static_cast<uint32_t>(dm.size());
    // many resident warps hide the decoder's latency: ask for the largest shared-memory carveout (23 KB of tables and output rings per CTA)
    cudaFuncSetAttribute(bgzf_inflate_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int inflate_ctas_per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&inflate_ctas_per_sm, bgzf_inflate_kernel, INF_WARPS * 32, 0);
    const uint32_t inflate_grid = std::min<uint32_t>((n_members + INF_WARPS - 1) / INF_WARPS, static_cast<uint32_t>(ctx->sm_count * std::max(inflate_ctas_per_sm, 1)));
    upload.worker.join();
    const double h2d_ms = upload.ms;
    if (upload.rc != SVB_OK) return fail(SVB_ERR_IO, "upload_file_range: short read or copy error");
    mark("upload");
    cudaEventRecord(ev[1], st);
    bgzf_inflate_kernel<<<inflate_grid, INF_WARPS * 32, 0, st>>>(d_comp, d_members, n_members, d_data, d_status);
    ctx->launches += 1;
    cudaEventRecord(ev[2], st);

    // while the GPU inflates: the record starts the index names (host work, hidden behind the inflate kernel)
    std::vector<uint64_t> index_starts;
    if (!getenv("SVB_INGEST_SERIAL_CHASE")) index_starts = bai_record_starts(std::string(path) + ".bai", members, 0, total_out);
    mark("index parsed");

    // ---- BAM header: host parse of the first bytes of the inflated stream
    std::unique_ptr<svb_bam> bam_owner(new (std::nothrow) svb_bam());
    svb_bam* bam = bam_owner.get();
    if (!bam) return fail(SVB_ERR_NOMEM, "svb_bam");
    uint64_t first_record = 0;
    {
        std::vector<uint8_t> head;
        uint64_t want = std::min<uint64_t>(total_out, 1u << 20);
        while (true) {
            head.resize(want);
            ING_CUDA(cudaMemcpyAsync(head.data(), d_data, want, cudaMemcpyDeviceToHost, st));
            uint32_t h_status = 0;
            ING_CUDA(cudaMemcpyAsync(&h_status, d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            ING_CUDA(cudaStreamSynchronize(st));
            if (h_status & ING_ERR_INFLATE) return fail(SVB_ERR_IO, "inflate failed (corrupt BGZF block)");
            std::string why;
            const int64_t used = bam_parse_header(head.data(), want, bam, &why);
            if (used >= 0) {
                first_record = static_cast<uint64_t>(used);
                break;
            }
            if (used == -1 || want == total_out) return fail(SVB_ERR_IO, why);
            want = std::min<uint64_t>(total_out, want * 8);
        }
    }
    const int32_t n_ref = static_cast<int32_t>(bam->contig_names.size());
    mark("inflate done, header parsed");

    // ---- record boundaries
    uint32_t cap = static_cast<uint32_t>(std::min<uint64_t>(0x7fffffffull, std::max<uint64_t>(1u << 16, (total_out - first_record) / 2048)));
    uint64_t* d_rec_off = nullptr;
    uint32_t n_rec = 0;
    cudaEventRecord(ev[3], st);
    bool chased = false;
    ctx->ingest_ms[11] = 0.0;
    {
        // record starts from the index: the walk becomes one short segment per thread (bam_chase_indexed_kernel)
        std::vector<uint64_t> starts;
        starts.reserve(index_starts.size() + 2);
        starts.push_back(first_record);
        for (uint64_t at : index_starts)
            if (at > first_record) starts.push_back(at);
        if (starts.size() >= 64) {
            starts.push_back(total_out);
            const uint32_t n_seg = static_cast<uint32_t>(starts.size() - 1);
            uint64_t* d_starts = nullptr;
            uint32_t* d_counts = nullptr;
            ING_CUDA(cudaMallocAsync(&d_starts, sizeof(uint64_t) * starts.size(), st));
            gc.dev.push_back(d_starts);
            ING_CUDA(cudaMallocAsync(&d_counts, sizeof(uint32_t) * (static_cast<size_t>(n_seg) + 2), st));
            gc.dev.push_back(d_counts);
            ING_CUDA(cudaMemcpyAsync(d_starts, starts.data(), sizeof(uint64_t) * starts.size(), cudaMemcpyHostToDevice, st));
            bam_chase_indexed_kernel<false><<<(n_seg + 127) / 128, 128, 0, st>>>(d_data, d_starts, n_seg, d_counts, nullptr, d_status);
            ctx->launches += 1;
            if (launch_scan_u32(ctx, d_counts, n_seg, ctx->d_counters + 15) != SVB_OK) return fail(SVB_ERR_CUDA, svb_last_error(ctx));
            uint32_t h2[4] = {0, 0, 0, 0};
            unsigned long long total = 0;
            ING_CUDA(cudaMemcpyAsync(h2, d_status, sizeof h2, cudaMemcpyDeviceToHost, st));
            ING_CUDA(cudaMemcpyAsync(&total, ctx->d_counters + 15, sizeof total, cudaMemcpyDeviceToHost, st));
            ING_CUDA(cudaStreamSynchronize(st));
            if (!(h2[0] & (ING_ERR_INDEX | ING_ERR_RECORD)) && total > 0 && total < 0x7fffffffull) {
                n_rec = static_cast<uint32_t>(total);
                inflate_cycles = static_cast<unsigned long long>(h2[2]) | (static_cast<unsigned long long>(h2[3]) << 32);
                ING_CUDA(cudaMallocAsync(&d_rec_off, sizeof(uint64_t) * n_rec, st));
                bam_chase_indexed_kernel<true><<<(n_seg + 127) / 128, 128, 0, st>>>(d_data, d_starts, n_seg, d_counts, d_rec_off, d_status);
                ctx->launches += 1;
                chased = true;
                ctx->ingest_ms[11] = static_cast<double>(n_seg);      // segments of the indexed walk (0: serial walk)
            } else {                     // the index does not describe this file: forget it
                ING_CUDA(cudaMemsetAsync(d_status, 0, sizeof(uint32_t), st));
            }
        }
    }
    for (int attempt = 0; attempt < 2 && !chased; ++attempt) {
        ING_CUDA(cudaMallocAsync(&d_rec_off, sizeof(uint64_t) * cap, st));
        bam_chase_kernel<<<1, 32, 0, st>>>(d_data, first_record, total_out, d_rec_off, cap, d_status + 1, d_status);
        ctx->launches += 1;
        uint32_t h2[4] = {0, 0, 0, 0};
        ING_CUDA(cudaMemcpyAsync(h2, d_status, sizeof h2, cudaMemcpyDeviceToHost, st));
        ING_CUDA(cudaStreamSynchronize(st));
        if (h2[0] & ING_ERR_RECORD) {
            cudaFreeAsync(d_rec_off, st);
            return fail(SVB_ERR_IO, "truncated BAM record");
        }
        n_rec = h2[1];
        inflate_cycles = static_cast<unsigned long long>(h2[2]) | (static_cast<unsigned long long>(h2[3]) << 32);
        if (n_rec <= cap) break;
        cudaFreeAsync(d_rec_off, st);
        d_rec_off = nullptr;
        cap = n_rec;
    }
    gc.dev.push_back(d_rec_off);
    cudaEventRecord(ev[4], st);
    mark("record chain");

    // ---- fixed fields + where the variable parts are
    svb_aln_hdr* d_hdr = nullptr;
    RecInfo* d_info = nullptr;
    ING_CUDA(cudaMallocAsync(&d_hdr, sizeof(svb_aln_hdr) * std::max<uint32_t>(n_rec, 1), st));
    gc.dev.push_back(d_hdr);
    ING_CUDA(cudaMallocAsync(&d_info, sizeof(RecInfo) * std::max<uint32_t>(n_rec, 1), st));
    gc.dev.push_back(d_info);
    bam->hdr.resize(n_rec);
    std::vector<RecInfo> info(n_rec);
    if (n_rec) {
        bam_fields_kernel<<<(n_rec + 127) / 128, 128, 0, st>>>(d_data, d_rec_off, n_rec, d_hdr, d_info, d_status);
        ctx->launches += 1;
        ING_CUDA(cudaMemcpyAsync(bam->hdr.data(), d_hdr, sizeof(svb_aln_hdr) * n_rec, cudaMemcpyDeviceToHost, st));
        ING_CUDA(cudaMemcpyAsync(info.data(), d_info, sizeof(RecInfo) * n_rec, cudaMemcpyDeviceToHost, st));
    }
    {
        uint32_t h_status = 0;
        ING_CUDA(cudaMemcpyAsync(&h_status, d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        ING_CUDA(cudaStreamSynchronize(st));
        if (h_status & ING_ERR_RECORD) return fail(SVB_ERR_IO, "malformed BAM record");
        if (h_status & ING_ERR_TAG) return fail(SVB_ERR_IO, "malformed auxiliary field");
    }

    mark("fixed fields");
    // ---- host prefix sums: where everything goes
    std::vector<CopyDst> dst(n_rec);
    bam->seq_off.resize(static_cast<size_t>(n_rec) + 1);
    bam->name_off.resize(static_cast<size_t>(n_rec) + 1);
    bam->sa_text_off.assign(n_rec, -1);
    uint64_t co = 0, so = 0, no = 0, sao = 0;
    for (uint32_t i = 0; i < n_rec; ++i) {
        const RecInfo& r = info[i];
        dst[i].cigar_op = co;
        dst[i].seq_byte = so;
        dst[i].name_byte = no;
        dst[i].sa_byte = sao;
        bam->hdr[i].cigar_off = co;
        bam->seq_off[i] = so;
        bam->name_off[i] = no;
        co += (static_cast<uint64_t>(r.n_cigar) + 3) / 4 * 4;
        so += (static_cast<uint64_t>(r.l_seq) + 1) / 2;
        no += r.name_len;
        if (r.sa_len != 0xFFFFFFFFu) {
            bam->sa_text_off[i] = static_cast<int64_t>(sao);
            sao += static_cast<uint64_t>(r.sa_len) + 1;
        }
    }
    bam->seq_off[n_rec] = so;
    bam->name_off[n_rec] = no;
    if (co / 4 >= 0xFFFFFFFFull) return fail(SVB_ERR_ARG, "more than 2^34 CIGAR ops");

    mark("host prefix sums");
    // ---- copies
    CopyDst* d_dst = nullptr;
    uint4* d_cigar = nullptr;
    uint8_t* d_seq4 = nullptr;
    char *d_names = nullptr, *d_sa = nullptr;
    ING_CUDA(cudaMallocAsync(&d_dst, sizeof(CopyDst) * std::max<uint32_t>(n_rec, 1), st));
    gc.dev.push_back(d_dst);
    ING_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_cigar), std::max<uint64_t>(cigar_padded_n4(co / 4), 1) * sizeof(uint4), st));
    const size_t gc_cigar = gc.dev.size();
    gc.dev.push_back(d_cigar);
    if (keep_sequences) ING_CUDA(cudaMallocAsync(&d_seq4, std::max<uint64_t>(so, 1), st));
    const size_t gc_seq = gc.dev.size();
    gc.dev.push_back(d_seq4);
    ING_CUDA(cudaMallocAsync(&d_names, std::max<uint64_t>(no, 1), st));
    gc.dev.push_back(d_names);
    ING_CUDA(cudaMallocAsync(&d_sa, std::max<uint64_t>(sao, 1), st));
    gc.dev.push_back(d_sa);
    bam->names.resize(no);
    bam->sa_text.resize(sao);
    if (n_rec) {
        ING_CUDA(cudaMemcpyAsync(d_dst, dst.data(), sizeof(CopyDst) * n_rec, cudaMemcpyHostToDevice, st));
        bam_copy_kernel<<<n_rec, 256, 0, st>>>(d_data, d_info, d_dst, n_rec, reinterpret_cast<uint32_t*>(d_cigar), d_seq4, d_names, d_sa);
        ctx->launches += 1;
        if (no) ING_CUDA(cudaMemcpyAsync(bam->names.data(), d_names, no, cudaMemcpyDeviceToHost, st));
        if (sao) ING_CUDA(cudaMemcpyAsync(bam->sa_text.data(), d_sa, sao, cudaMemcpyDeviceToHost, st));
    }
    cudaEventRecord(ev[5], st);
    {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail(SVB_ERR_CUDA, std::string("device ingest: ") + cudaGetErrorString(e));
    }
    const auto wall_dev = std::chrono::steady_clock::now();
    mark("copies");

    // ---- SA tags: the same host parser as the host ingest (retrieve_other_alignments, SVIM_COLLECT.py:8-58)
    std::vector<const char*> name_ptrs;
    for (auto& s : bam->contig_names) name_ptrs.push_back(s.c_str());
    bam->sa_count.assign(n_rec, 0);
    for (uint32_t i = 0; i < n_rec; ++i) {
        svb_aln_hdr& h = bam->hdr[i];
        h.sa_first = static_cast<uint32_t>(bam->seg.size());
        if (bam->sa_text_off[i] < 0) continue;
        const char* sa = bam->sa_text.data() + bam->sa_text_off[i];
        const int cnt = svb_parse_sa(sa, name_ptrs.data(), n_ref, nullptr, 0);
        if (cnt < 0) return fail(cnt, std::string("malformed SA tag (int() would raise): ") + sa);
        bam->seg.resize(bam->seg.size() + static_cast<size_t>(cnt));
        svb_parse_sa(sa, name_ptrs.data(), n_ref, bam->seg.data() + h.sa_first, cnt);
        bam->sa_count[i] = static_cast<uint32_t>(cnt);
    }

    mark("SA tags");
    // ---- the record image (python string order of the contig names is the caller's: SVCandidate.py:352)
    std::vector<int32_t> lexrank(static_cast<size_t>(std::max(n_ref, 0)));
    if (contig_lexrank_or_null) {
        std::copy(contig_lexrank_or_null, contig_lexrank_or_null + n_ref, lexrank.begin());
    } else {                          // code-point order of the (ASCII) names == python's str order for them
        std::vector<int32_t> order(lexrank.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = static_cast<int32_t>(i);
        std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return bam->contig_names[a] < bam->contig_names[b]; });
        for (size_t i = 0; i < order.size(); ++i) lexrank[static_cast<size_t>(order[i])] = static_cast<int32_t>(i);
    }
    svb_records* rec = nullptr;
    gc.dev[gc_cigar] = nullptr;       // the records own d_cigar from the call on, also when it fails
    int rc = load_records_impl(ctx, bam->hdr.data(), n_rec, nullptr, d_cigar, co, bam->seg.data(), bam->sa_count.data(),
                               static_cast<uint32_t>(bam->seg.size()), bam->contig_len.data(), lexrank.data(), n_ref, &rec);
    if (rc != SVB_OK) {
        cudaStreamSynchronize(st);
        for (auto& e : ev) cudaEventDestroy(e);
        set_err(err, err_len, svb_last_error(ctx));
        return rc;
    }
    if (keep_sequences) {
        gc.dev[gc_seq] = nullptr;
        rec->d_seq4 = d_seq4;
        rec->seq_bytes = so;
        cudaError_t e = cudaMallocAsync(&rec->d_seq_off, sizeof(uint64_t) * (static_cast<size_t>(n_rec) + 1), st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(rec->d_seq_off, bam->seq_off.data(), sizeof(uint64_t) * (static_cast<size_t>(n_rec) + 1), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            svb_records_free(rec);
            return fail(SVB_ERR_CUDA, std::string("device ingest sequences: ") + cudaGetErrorString(e));
        }
    }
    const auto wall_end = std::chrono::steady_clock::now();
    mark("record image");
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    float f = 0.f;
    ctx->ingest_ms[0] = ms(wall0, wall_read);
    ctx->ingest_ms[1] = h2d_ms;                 // host wall clock: parallel pread + staged copies (file_upload.cu)
    cudaEventElapsedTime(&f, ev[1], ev[2]); ctx->ingest_ms[2] = f;
    cudaEventElapsedTime(&f, ev[3], ev[4]); ctx->ingest_ms[3] = f;
    cudaEventElapsedTime(&f, ev[4], ev[5]); ctx->ingest_ms[4] = f;
    ctx->ingest_ms[5] = ms(wall_dev, wall_end);
    ctx->ingest_ms[6] = ms(wall0, wall_end);
    ctx->ingest_ms[7] = static_cast<double>(total_out);
    ctx->ingest_ms[8] = inflate_ctas_per_sm;
    ctx->ingest_ms[9] = n_members ? static_cast<double>(inflate_cycles) / n_members : 0.0;
    ctx->ingest_ms[10] = n_members;
    for (auto& e : ev) cudaEventDestroy(e);
#undef ING_CUDA
    *bam_out = bam_owner.release();
    *rec_out = rec;
    return SVB_OK;
}

// the host copies the per-alignment seams read (cigartuples, query_sequence) are NOT made by the device ingest;
// this downloads them on demand
int svb_bam_materialize_host(svb_ctx* ctx, svb_bam* bam, const svb_records* rec, int what) {
    if (!ctx || !bam || !rec) return ctx ? svb_fail(ctx, SVB_ERR_ARG, "svb_bam_materialize_host") : SVB_ERR_ARG;
    cudaSetDevice(ctx->device);
    const uint64_t n_ops = rec->n4 * 4;
    if ((what & 1) && bam->cigar.size() != n_ops) {
        bam->cigar.resize(n_ops);
        if (n_ops) SVB_CUDA(ctx, cudaMemcpyAsync(bam->cigar.data(), rec->d_cigar, sizeof(uint32_t) * n_ops, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if ((what & 2) && rec->d_seq4 && bam->seq4.size() != rec->seq_bytes) {
        bam->seq4.resize(rec->seq_bytes);
        if (rec->seq_bytes) SVB_CUDA(ctx, cudaMemcpyAsync(bam->seq4.data(), rec->d_seq4, rec->seq_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SVB_OK;
}

}  // extern "C"
