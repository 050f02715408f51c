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
#include <cstring>
#include <memory>
#include <new>
#include <string>
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
    ING_ERR_TAG = 4u            // malformed auxiliary field
};

struct DevMember {
    uint64_t in_off, out_off;
    uint32_t in_len, out_len;
};

// One WARP per member.  A thread per member would put 32 unrelated decoder states into one warp: they diverge at every
// branch and the warp runs them one after the other (measured: 2.5 GB/s).  Here lane 0 decodes (Huffman tables in shared
// memory, canonical count/symbol form of inflate_core.cuh) and writes the literals; at every match and stored block all
// 32 lanes copy together; the other lanes wait at the broadcast.  Members of one CTA never interact.
constexpr int INF_WARPS = 4;
constexpr int FAST_LEN_BITS = 10, FAST_DIST_BITS = 9;
struct InfTables {
    uint16_t lencnt[16], lensym[288], distcnt[16], distsym[32];
    uint8_t lengths[320];
    // direct lookup on the next bits of the stream: (symbol << 4) | code length, 0 = code longer than the table (slow path)
    uint16_t fast_len[1 << FAST_LEN_BITS], fast_dist[1 << FAST_DIST_BITS];
};

// all lanes: fill the direct table of a canonical code (count / symbol form).  Deflate sends codes most significant
// bit first into a stream that is read least significant bit first, hence the bit reversal.
__device__ void build_fast_table(const uint16_t* count, const uint16_t* symbol, uint16_t* table, int bits, uint32_t lane) {
    const uint32_t size = 1u << bits;
    for (uint32_t i = lane; i < size; i += 32u) table[i] = 0;
    __syncwarp();
    uint32_t total = 0;
    for (int len = 1; len <= 15; ++len) total += count[len];
    for (uint32_t j = lane; j < total; j += 32u) {
        uint32_t code = 0, first_idx = 0;
        int L = 0;
        for (int len = 1; len <= 15; ++len) {
            const uint32_t cnt = count[len];
            if (j < first_idx + cnt) {
                L = len;
                code += j - first_idx;
                break;
            }
            code = (code + cnt) << 1;
            first_idx += cnt;
        }
        if (L == 0 || L > bits) continue;
        const uint32_t rev = __brev(code) >> (32 - L);
        const uint16_t entry = static_cast<uint16_t>((static_cast<uint32_t>(symbol[j]) << 4) | static_cast<uint32_t>(L));
        for (uint32_t k = rev; k < size; k += 1u << L) table[k] = entry;
    }
    __syncwarp();
}

__device__ __forceinline__ int decode_fast(InfBits& b, const InfHuff& h, const uint16_t* table, int bits) {
    inf_need(b, 15);
    const uint32_t e = table[static_cast<uint32_t>(b.buf) & ((1u << bits) - 1u)];
    if (e) {
        b.buf >>= (e & 15u);
        b.cnt -= static_cast<int>(e & 15u);
        return static_cast<int>(e >> 4);
    }
    return inf_decode(b, h);
}

__device__ int inflate_member_warp(const uint8_t* __restrict__ src, uint32_t src_len, uint8_t* dst, uint32_t out_len, InfTables& T,
                                   uint32_t lane) {
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    constexpr uint32_t FULLM = 0xffffffffu;
    InfHuff lencode{T.lencnt, T.lensym}, distcode{T.distcnt, T.distsym};
    InfBits b{src, src + src_len, 0ull, 0, 0};          // lane 0's
    uint32_t pos = 0;                                   // warp-uniform
    int last = 0;
    do {
        // ---- block header and code tables: lane 0
        int err = 0;
        uint32_t type = 0, st_off = 0, st_len = 0;
        if (lane == 0) {
            last = static_cast<int>(inf_bits(b, 1));
            type = inf_bits(b, 2);
            if (type == 0u) {                            // stored: where the bytes are
                if (inf_overrun(b)) err = INF_ERR_INPUT;
                else {
                    b.p -= (b.cnt - b.virt) >> 3;
                    b.buf = 0; b.cnt = 0; b.virt = 0;
                    if (b.end - b.p < 4) err = INF_ERR_INPUT;
                    else {
                        const uint32_t len = b.p[0] | (static_cast<uint32_t>(b.p[1]) << 8);
                        const uint32_t nlen = b.p[2] | (static_cast<uint32_t>(b.p[3]) << 8);
                        b.p += 4;
                        if (len != (~nlen & 0xFFFFu)) err = INF_ERR_CODE;
                        else if (static_cast<uint32_t>(b.end - b.p) < len) err = INF_ERR_INPUT;
                        else if (pos + len > out_len) err = INF_ERR_OUTPUT;
                        else {
                            st_off = static_cast<uint32_t>(b.p - src);
                            st_len = len;
                            b.p += len;
                        }
                    }
                }
            } else if (type == 1u) {
                int s = 0;
                for (; s < 144; ++s) T.lengths[s] = 8;
                for (; s < 256; ++s) T.lengths[s] = 9;
                for (; s < 280; ++s) T.lengths[s] = 7;
                for (; s < 288; ++s) T.lengths[s] = 8;
                inf_construct(lencode, T.lengths, 288);
                for (s = 0; s < 30; ++s) T.lengths[s] = 5;
                inf_construct(distcode, T.lengths, 30);
            } else if (type == 2u) {
                const int nlen = static_cast<int>(inf_bits(b, 5)) + 257;
                const int ndist = static_cast<int>(inf_bits(b, 5)) + 1;
                const int ncode = static_cast<int>(inf_bits(b, 4)) + 4;
                if (nlen > 286 || ndist > 30) err = INF_ERR_CODE;
                else {
                    int idx = 0;
                    for (; idx < ncode; ++idx) T.lengths[order[idx]] = static_cast<uint8_t>(inf_bits(b, 3));
                    for (; idx < 19; ++idx) T.lengths[order[idx]] = 0;
                    if (inf_construct(lencode, T.lengths, 19) != 0) err = INF_ERR_CODE;
                    idx = 0;
                    while (!err && idx < nlen + ndist) {
                        const int sym = inf_decode(b, lencode);
                        if (sym < 0) { err = INF_ERR_CODE; break; }
                        if (sym < 16) {
                            T.lengths[idx++] = static_cast<uint8_t>(sym);
                        } else {
                            int len = 0, rep;
                            if (sym == 16) {
                                if (idx == 0) { err = INF_ERR_CODE; break; }
                                len = T.lengths[idx - 1];
                                rep = 3 + static_cast<int>(inf_bits(b, 2));
                            } else if (sym == 17) {
                                rep = 3 + static_cast<int>(inf_bits(b, 3));
                            } else {
                                rep = 11 + static_cast<int>(inf_bits(b, 7));
                            }
                            if (idx + rep > nlen + ndist) { err = INF_ERR_CODE; break; }
                            while (rep--) T.lengths[idx++] = static_cast<uint8_t>(len);
                        }
                    }
                    if (!err && T.lengths[256] == 0) err = INF_ERR_CODE;
                    if (!err) {
                        // the distance lengths move out of the way first: constructing the length code reuses nothing of them
                        int e1 = inf_construct(distcode, T.lengths + nlen, ndist);
                        if (e1 < 0 || (e1 > 0 && ndist - distcode.count[0] != 1)) err = INF_ERR_CODE;
                        e1 = inf_construct(lencode, T.lengths, nlen);
                        if (e1 < 0 || (e1 > 0 && nlen - lencode.count[0] != 1)) err = INF_ERR_CODE;
                    }
                }
            } else {
                err = INF_ERR_CODE;
            }
        }
        err = __shfl_sync(FULLM, err, 0);
        if (err) return err;
        type = __shfl_sync(FULLM, type, 0);
        last = __shfl_sync(FULLM, last, 0);
        if (type != 0u) {                               // (the shuffles above made lane 0's tables visible)
            __syncwarp();
            build_fast_table(T.lencnt, T.lensym, T.fast_len, FAST_LEN_BITS, lane);
            build_fast_table(T.distcnt, T.distsym, T.fast_dist, FAST_DIST_BITS, lane);
        }
        if (type == 0u) {                               // stored block: all lanes copy
            st_off = __shfl_sync(FULLM, st_off, 0);
            st_len = __shfl_sync(FULLM, st_len, 0);
            for (uint32_t i = lane; i < st_len; i += 32u) dst[pos + i] = src[st_off + i];
            pos += st_len;
            __syncwarp();
            continue;
        }
        // ---- symbols: lane 0 writes literals until it meets a match or the end of the block
        while (true) {
            uint32_t ev_len = 0, ev_dist = 0;           // ev_len == 0: end of block
            if (lane == 0) {
                while (true) {
                    int sym = decode_fast(b, lencode, T.fast_len, FAST_LEN_BITS);
                    if (sym < 0) { err = INF_ERR_CODE; break; }
                    if (sym < 256) {
                        if (pos >= out_len) { err = INF_ERR_OUTPUT; break; }
                        dst[pos++] = static_cast<uint8_t>(sym);
                        if (inf_overrun(b)) { err = INF_ERR_INPUT; break; }
                        continue;
                    }
                    if (sym == 256) break;
                    sym -= 257;
                    if (sym >= 29) { err = INF_ERR_CODE; break; }
                    ev_len = lbase[sym] + inf_bits(b, lext[sym]);
                    const int dsym = decode_fast(b, distcode, T.fast_dist, FAST_DIST_BITS);
                    if (dsym < 0 || dsym >= 30) { err = INF_ERR_CODE; break; }
                    ev_dist = dbase[dsym] + inf_bits(b, dext[dsym]);
                    if (ev_dist > pos) err = INF_ERR_CODE;
                    else if (pos + ev_len > out_len) err = INF_ERR_OUTPUT;
                    break;
                }
                if (!err && inf_overrun(b)) err = INF_ERR_INPUT;
            }
            err = __shfl_sync(FULLM, err, 0);           // (also orders lane 0's literal stores before the copy below)
            if (err) return err;
            pos = __shfl_sync(FULLM, pos, 0);
            ev_len = __shfl_sync(FULLM, ev_len, 0);
            if (ev_len == 0u) break;
            ev_dist = __shfl_sync(FULLM, ev_dist, 0);
            __syncwarp();
            const uint8_t* from = dst + pos - ev_dist;
            if (ev_dist >= ev_len) {
                for (uint32_t i = lane; i < ev_len; i += 32u) dst[pos + i] = from[i];
            } else {                                    // overlapping copy = the last ev_dist bytes repeated
                for (uint32_t i = lane; i < ev_len; i += 32u) dst[pos + i] = from[i % ev_dist];
            }
            __syncwarp();
            pos += ev_len;
        }
    } while (!last);
    return pos == out_len ? INF_OK : INF_ERR_SIZE;
}

__global__ void __launch_bounds__(INF_WARPS * 32) bgzf_inflate_kernel(const uint8_t* __restrict__ comp, const DevMember* __restrict__ members,
                                                                      uint32_t n, uint8_t* out, uint32_t* status) {
    __shared__ InfTables tables[INF_WARPS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t i = blockIdx.x * INF_WARPS + warp;
    if (i >= n) return;
    const DevMember m = members[i];
    const int rc = inflate_member_warp(comp + m.in_off, m.in_len, out + m.out_off, m.out_len, tables[warp], lane);
    if (rc != INF_OK && lane == 0) atomicOr(status, ING_ERR_INFLATE);
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

// timings of the last svb_bam_open_device call (ms): read file, H2D, inflate, chase, fields + copy, host parse, total
static double g_ingest_ms[8] = {0};
const double* svb_bam_device_timings(void) { return g_ingest_ms; }

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

    std::vector<BgzfMember> members;
    uint64_t total_out = 0;
    {
        std::string why;
        if (!bgzf_member_table(raw, static_cast<uint64_t>(fsize), &members, &total_out, &why)) return fail(SVB_ERR_IO, why);
    }
    if (members.empty() || members.size() > 0x7fffffffull) return fail(SVB_ERR_IO, "not a BAM file");
    std::vector<DevMember> dm(members.size());
    for (size_t i = 0; i < members.size(); ++i) {
        dm[i].in_off = members[i].in_off;
        dm[i].out_off = members[i].out_off;
        dm[i].in_len = static_cast<uint32_t>(members[i].in_len);
        dm[i].out_len = static_cast<uint32_t>(members[i].out_len);
    }

    // ---- H2D + inflate
    uint8_t *d_comp = nullptr, *d_data = nullptr;
    DevMember* d_members = nullptr;
    uint32_t* d_status = nullptr;      // [0] error bits, [1] record count
#define ING_CUDA(call)                                                             \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) return fail(SVB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)
    ING_CUDA(cudaMallocAsync(&d_comp, static_cast<size_t>(fsize), st));
    gc.dev.push_back(d_comp);
    ING_CUDA(cudaMallocAsync(&d_data, std::max<uint64_t>(total_out, 1), st));
    gc.dev.push_back(d_data);
    ING_CUDA(cudaMallocAsync(&d_members, sizeof(DevMember) * dm.size(), st));
    gc.dev.push_back(d_members);
    ING_CUDA(cudaMallocAsync(&d_status, 2 * sizeof(uint32_t), st));
    gc.dev.push_back(d_status);
    ING_CUDA(cudaMemsetAsync(d_status, 0, 2 * sizeof(uint32_t), st));
    cudaEventRecord(ev[0], st);
    const auto wall_h2d = std::chrono::steady_clock::now();
    if (upload_file_range(ctx, fd, 0, static_cast<uint64_t>(fsize), d_comp) != SVB_OK) return fail(SVB_ERR_IO, svb_last_error(ctx));
    const double h2d_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall_h2d).count();
    ING_CUDA(cudaMemcpyAsync(d_members, dm.data(), sizeof(DevMember) * dm.size(), cudaMemcpyHostToDevice, st));
    cudaEventRecord(ev[1], st);
    const uint32_t n_members = static_cast<uint32_t>(dm.size());
    bgzf_inflate_kernel<<<(n_members + INF_WARPS - 1) / INF_WARPS, INF_WARPS * 32, 0, st>>>(d_comp, d_members, n_members, d_data, d_status);
    ctx->launches += 1;
    cudaEventRecord(ev[2], st);

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

    // ---- record boundaries
    uint32_t cap = static_cast<uint32_t>(std::min<uint64_t>(0x7fffffffull, std::max<uint64_t>(1u << 16, (total_out - first_record) / 2048)));
    uint64_t* d_rec_off = nullptr;
    uint32_t n_rec = 0;
    cudaEventRecord(ev[3], st);
    for (int attempt = 0; attempt < 2; ++attempt) {
        ING_CUDA(cudaMallocAsync(&d_rec_off, sizeof(uint64_t) * cap, st));
        bam_chase_kernel<<<1, 32, 0, st>>>(d_data, first_record, total_out, d_rec_off, cap, d_status + 1, d_status);
        ctx->launches += 1;
        uint32_t h2[2] = {0, 0};
        ING_CUDA(cudaMemcpyAsync(h2, d_status, sizeof h2, cudaMemcpyDeviceToHost, st));
        ING_CUDA(cudaStreamSynchronize(st));
        if (h2[0] & ING_ERR_RECORD) {
            cudaFreeAsync(d_rec_off, st);
            return fail(SVB_ERR_IO, "truncated BAM record");
        }
        n_rec = h2[1];
        if (n_rec <= cap) break;
        cudaFreeAsync(d_rec_off, st);
        d_rec_off = nullptr;
        cap = n_rec;
    }
    gc.dev.push_back(d_rec_off);
    cudaEventRecord(ev[4], st);

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
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    float f = 0.f;
    g_ingest_ms[0] = ms(wall0, wall_read);
    g_ingest_ms[1] = h2d_ms;                 // host wall clock: parallel pread + staged copies (file_upload.cu)
    cudaEventElapsedTime(&f, ev[1], ev[2]); g_ingest_ms[2] = f;
    cudaEventElapsedTime(&f, ev[3], ev[4]); g_ingest_ms[3] = f;
    cudaEventElapsedTime(&f, ev[4], ev[5]); g_ingest_ms[4] = f;
    g_ingest_ms[5] = ms(wall_dev, wall_end);
    g_ingest_ms[6] = ms(wall0, wall_end);
    g_ingest_ms[7] = static_cast<double>(total_out);
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
