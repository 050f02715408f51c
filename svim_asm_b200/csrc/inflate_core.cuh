// DEFLATE (RFC 1951) decoder for ONE BGZF member.
//
// Where it sits: the reference opens its input with pysam.AlignmentFile (svim-asm:63,85-86), i.e. htslib inflates
// the BGZF blocks of the BAM file on the host [ext].  BGZF members are independent raw-deflate streams of at most
// 64 KiB each, so a whole-genome BAM is tens of thousands of independent jobs.  Two decoders live here:
//   * inflate_member: the plain one (canonical-Huffman count/symbol arrays, byte-wise I/O), the specification the other is
//     tested against and the slow path for codes longer than the direct tables;
//   * the pieces of the device decoder (bam_device.cu: one WARP per member): block header, direct tables, the symbol loop
//     that every lane of the warp runs in lock step (inf_run_lanes), the output ring and its flush.
// The same code compiles for the host (tests/hostcheck) where both are checked against zlib.
#pragma once
#include <stdint.h>

#include "linkage.cuh"   // SVB_HD

enum : int {
    INF_OK = 0,
    INF_ERR_INPUT = 1,      // ran out of compressed bytes
    INF_ERR_OUTPUT = 2,     // more output than the member's ISIZE
    INF_ERR_CODE = 3,       // invalid block type / code lengths / symbol / distance
    INF_ERR_SIZE = 4        // fewer output bytes than ISIZE
};

struct InfBits {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t buf;
    int cnt;
    int virt;               // zero bits appended past the end of the input (look-ahead may go there, consumption may not)
};
SVB_HD bool inf_overrun(const InfBits& b) { return b.cnt < b.virt; }

SVB_HD void inf_need(InfBits& b, int n) {
    while (b.cnt < n) {
        uint64_t byte = 0;
        if (b.p < b.end) byte = *b.p++;
        else b.virt += 8;
        b.buf |= byte << b.cnt;
        b.cnt += 8;
    }
}
SVB_HD uint32_t inf_bits(InfBits& b, int n) {           // n <= 16
    if (n == 0) return 0u;
    inf_need(b, n);
    const uint32_t v = static_cast<uint32_t>(b.buf) & ((1u << n) - 1u);
    b.buf >>= n;
    b.cnt -= n;
    return v;
}

// canonical Huffman code: count[len] codes of each length, symbols in code order
struct InfHuff {
    uint16_t* count;        // [16]
    uint16_t* symbol;       // [n]
};

// returns 0 for a complete code, > 0 for an incomplete one, < 0 for an over-subscribed one
SVB_HD int inf_construct(InfHuff& h, const uint8_t* length, int n) {
    for (int len = 0; len <= 15; ++len) h.count[len] = 0;
    for (int s = 0; s < n; ++s) h.count[length[s]]++;
    if (h.count[0] == n) return 0;
    int left = 1;
    for (int len = 1; len <= 15; ++len) {
        left <<= 1;
        left -= h.count[len];
        if (left < 0) return left;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int len = 1; len < 15; ++len) offs[len + 1] = static_cast<uint16_t>(offs[len] + h.count[len]);
    for (int s = 0; s < n; ++s)
        if (length[s] != 0) h.symbol[offs[length[s]]++] = static_cast<uint16_t>(s);
    return left;
}

// one symbol, bit by bit (at most 15 rounds); -1 on an invalid code
SVB_HD int inf_decode(InfBits& b, const InfHuff& h) {
    inf_need(b, 15);
    uint32_t bitbuf = static_cast<uint32_t>(b.buf);
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        code |= static_cast<int>(bitbuf & 1u);
        bitbuf >>= 1;
        const int count = h.count[len];
        if (code - count < first) {
            b.buf >>= len;
            b.cnt -= len;
            return h.symbol[index + (code - first)];
        }
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// Inflate one raw-deflate stream of exactly `out_len` bytes.
SVB_HD int inflate_member(const uint8_t* src, uint32_t src_len, uint8_t* dst, uint32_t out_len) {
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

    uint16_t lencnt[16], lensym[288], distcnt[16], distsym[30];
    InfHuff lencode{lencnt, lensym}, distcode{distcnt, distsym};
    uint8_t lengths[320];
    InfBits b{src, src + src_len, 0ull, 0, 0};
    uint32_t pos = 0;
    int last = 0;
    do {
        last = static_cast<int>(inf_bits(b, 1));
        const uint32_t type = inf_bits(b, 2);
        if (type == 0u) {                                   // stored
            if (inf_overrun(b)) return INF_ERR_INPUT;
            b.p -= (b.cnt - b.virt) >> 3;                  // give back whole bytes of look-ahead, drop the rest of the current byte
            b.buf = 0;
            b.cnt = 0;
            b.virt = 0;
            if (b.end - b.p < 4) return INF_ERR_INPUT;
            const uint32_t len = b.p[0] | (static_cast<uint32_t>(b.p[1]) << 8);
            const uint32_t nlen = b.p[2] | (static_cast<uint32_t>(b.p[3]) << 8);
            b.p += 4;
            if (len != (~nlen & 0xFFFFu)) return INF_ERR_CODE;
            if (static_cast<uint32_t>(b.end - b.p) < len) return INF_ERR_INPUT;
            if (pos + len > out_len) return INF_ERR_OUTPUT;
            for (uint32_t i = 0; i < len; ++i) dst[pos + i] = b.p[i];
            b.p += len;
            pos += len;
            continue;
        }
        if (type == 1u) {                                   // fixed codes
            int s = 0;
            for (; s < 144; ++s) lengths[s] = 8;
            for (; s < 256; ++s) lengths[s] = 9;
            for (; s < 280; ++s) lengths[s] = 7;
            for (; s < 288; ++s) lengths[s] = 8;
            inf_construct(lencode, lengths, 288);
            for (s = 0; s < 30; ++s) lengths[s] = 5;
            inf_construct(distcode, lengths, 30);
        } else if (type == 2u) {                            // dynamic codes
            const int nlen = static_cast<int>(inf_bits(b, 5)) + 257;
            const int ndist = static_cast<int>(inf_bits(b, 5)) + 1;
            const int ncode = static_cast<int>(inf_bits(b, 4)) + 4;
            if (nlen > 286 || ndist > 30) return INF_ERR_CODE;
            int idx = 0;
            for (; idx < ncode; ++idx) lengths[order[idx]] = static_cast<uint8_t>(inf_bits(b, 3));
            for (; idx < 19; ++idx) lengths[order[idx]] = 0;
            if (inf_construct(lencode, lengths, 19) != 0) return INF_ERR_CODE;      // the code-length code must be complete
            idx = 0;
            while (idx < nlen + ndist) {
                int sym = inf_decode(b, lencode);
                if (sym < 0) return INF_ERR_CODE;
                if (sym < 16) {
                    lengths[idx++] = static_cast<uint8_t>(sym);
                } else {
                    int len = 0, rep;
                    if (sym == 16) {
                        if (idx == 0) return INF_ERR_CODE;
                        len = lengths[idx - 1];
                        rep = 3 + static_cast<int>(inf_bits(b, 2));
                    } else if (sym == 17) {
                        rep = 3 + static_cast<int>(inf_bits(b, 3));
                    } else {
                        rep = 11 + static_cast<int>(inf_bits(b, 7));
                    }
                    if (idx + rep > nlen + ndist) return INF_ERR_CODE;
                    while (rep--) lengths[idx++] = static_cast<uint8_t>(len);
                }
            }
            if (lengths[256] == 0) return INF_ERR_CODE;     // no end-of-block code
            int err = inf_construct(lencode, lengths, nlen);
            if (err < 0 || (err > 0 && nlen - lencode.count[0] != 1)) return INF_ERR_CODE;
            err = inf_construct(distcode, lengths + nlen, ndist);
            if (err < 0 || (err > 0 && ndist - distcode.count[0] != 1)) return INF_ERR_CODE;
        } else {
            return INF_ERR_CODE;
        }
        // literal / length + distance symbols until end-of-block; every round emits at least one byte or ends the block
        while (true) {
            int sym = inf_decode(b, lencode);
            if (sym < 0) return INF_ERR_CODE;
            if (sym < 256) {
                if (pos >= out_len) return INF_ERR_OUTPUT;
                dst[pos++] = static_cast<uint8_t>(sym);
            } else if (sym == 256) {
                break;
            } else {
                sym -= 257;
                if (sym >= 29) return INF_ERR_CODE;
                const uint32_t len = lbase[sym] + inf_bits(b, lext[sym]);
                const int dsym = inf_decode(b, distcode);
                if (dsym < 0 || dsym >= 30) return INF_ERR_CODE;
                const uint32_t dist = dbase[dsym] + inf_bits(b, dext[dsym]);
                if (dist > pos) return INF_ERR_CODE;
                if (pos + len > out_len) return INF_ERR_OUTPUT;
                for (uint32_t i = 0; i < len; ++i, ++pos) dst[pos] = dst[pos - dist];
            }
            if (inf_overrun(b)) return INF_ERR_INPUT;
        }
        if (inf_overrun(b)) return INF_ERR_INPUT;
    } while (!last);
    return pos == out_len ? INF_OK : INF_ERR_SIZE;
}

// ---- the fast symbol loop of the device decoder (bam_device.cu; one WARP per member, every lane runs this loop in lock step) ----
// What counts is instructions per symbol and the latency of the chain from one symbol to the next:
//   * the bit buffer is refilled 32 bits at a time from aligned words (bytewise only at a member's unaligned head and tail),
//   * the direct tables hold finished entries -- literal value, or length / distance BASE with the number of extra bits --
//     so that a symbol is one table load plus shifts,
//   * matches are copied by the lanes together, a byte each (98 % of those in BAM data are at most 16 bytes: 4-byte CIGAR
//     words, short repeats).
constexpr int INF_LEN_BITS = 10, INF_DIST_BITS = 8;
// Literal / length table entry, 16 bits (the table is the largest piece of a warp's shared memory, and the number of resident
// warps is what hides the decoder's latency): bits 0-3 code length (0 = code longer than the table: slow path), bit 4 set for
// a length symbol, bits 5-7 its number of extra bits (7 = special: value 0 end of block, value 1 invalid symbol), bits 8-15 the
// literal byte, or the length base - 3.
typedef uint16_t inf_len_t;
enum : uint32_t { INF_L_LEN = 1u << 4, INF_L_SPECIAL = 7u << 5 };
// Distance table entry, 32 bits: bits 0-3 code length (0 = slow path), 4-5 kind, 8-11 extra bits, 16-31 distance base.
enum : uint32_t { INF_K_LIT = 0u, INF_K_LEN = 1u << 4, INF_K_EOB = 2u << 4, INF_K_BAD = 3u << 4, INF_K_MASK = 3u << 4 };

SVB_HD uint32_t inf_len_entry(uint32_t sym, uint32_t code_len) {
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    if (sym < 256u) return (sym << 8) | code_len;
    if (sym == 256u) return INF_L_LEN | INF_L_SPECIAL | code_len;
    if (sym >= 286u) return (1u << 8) | INF_L_LEN | INF_L_SPECIAL | code_len;
    return (static_cast<uint32_t>(lbase[sym - 257u] - 3u) << 8) | (static_cast<uint32_t>(lext[sym - 257u]) << 5) | INF_L_LEN | code_len;
}
SVB_HD uint32_t inf_dist_entry(uint32_t sym, uint32_t code_len) {
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    if (sym >= 30u) return INF_K_BAD | code_len;
    return (static_cast<uint32_t>(dbase[sym]) << 16) | (static_cast<uint32_t>(dext[sym]) << 8) | INF_K_LEN | code_len;
}

SVB_HD uint32_t inf_bitrev(uint32_t v, int n) {          // the low n bits of v, reversed
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

// Direct table of a canonical code: entry for every `bits`-bit window whose leading bits are a code of length <= bits.
// Lanes split the symbols (lane, n_lanes); the caller has zeroed the table and synchronises afterwards.  Deflate packs codes most
// significant bit first into a stream read least significant bit first, hence the reversal.
template <typename Entry>
SVB_HD void inf_fill_table(const uint16_t* count, const uint16_t* symbol, Entry* table, int bits, bool dist, uint32_t lane,
                           uint32_t n_lanes) {
    const uint32_t size = 1u << bits;
    uint32_t total = 0;
    for (int len = 1; len <= 15; ++len) total += count[len];
    for (uint32_t j = lane; j < total; j += n_lanes) {
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
        const Entry entry = static_cast<Entry>(dist ? inf_dist_entry(symbol[j], static_cast<uint32_t>(L)) : inf_len_entry(symbol[j], static_cast<uint32_t>(L)));
        for (uint32_t k = inf_bitrev(code, L); k < size; k += 1u << L) table[k] = entry;
    }
}

// at least 32 bits in the buffer afterwards (zeros past the end of the input, counted in virt)
SVB_HD void inf_refill(InfBits& b) {
    if ((reinterpret_cast<uintptr_t>(b.p) & 3u) == 0u && b.end - b.p >= 4) {
#ifdef __CUDA_ARCH__
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(b.p));
#else
        const uint32_t w = static_cast<uint32_t>(b.p[0]) | (static_cast<uint32_t>(b.p[1]) << 8) | (static_cast<uint32_t>(b.p[2]) << 16) |
                           (static_cast<uint32_t>(b.p[3]) << 24);
#endif
        b.buf |= static_cast<uint64_t>(w) << b.cnt;              // cnt < 32 here
        b.p += 4;
        b.cnt += 32;
        return;
    }
    // bytewise: the unaligned head of a member (on until the pointer is aligned), its last bytes, virtual zeros beyond
    while (b.cnt < 32 || ((reinterpret_cast<uintptr_t>(b.p) & 3u) != 0u && b.cnt <= 56 && b.p < b.end)) {
        uint64_t byte = 0;
        if (b.p < b.end) byte = *b.p++;
        else b.virt += 8;
        b.buf |= byte << b.cnt;
        b.cnt += 8;
    }
}

// ---- output window -------------------------------------------------------------------------------------------------------
// The decoder writes into a ring of INF_RING bytes (shared memory on the device) that is both the staging buffer for the
// output -- flushed to the member's place in global memory in 16-byte vectors by the whole warp -- and the window recent
// matches are copied from.  Why: with bytes stored one by one to global memory and matches reading them back, every
// match is a round trip to L2 (or to DRAM: the lines of all resident members do not fit L2) and every literal a partial
// sector write; measured 850 clock cycles per symbol.  With the ring, the symbol loop only touches shared memory and the
// compressed input.  Ring index of output byte p = (address of that byte in global memory) mod INF_RING, so ring and
// global memory agree on 16-byte alignment.  A match whose source is no longer in the ring (distance + length >
// INF_RING) reads global memory: that range has been flushed, because at most INF_FLUSH_AT + 258 bytes are ever unflushed.
// (Measured on B200: the symbol loop is bound by dependent-instruction latency, about 5 cycles per instruction per warp, so
// the sizes are chosen for 7 CTAs = 28 decoding warps per SM rather than for the largest window.)
constexpr uint32_t INF_RING = 2048, INF_FLUSH_AT = 1024, INF_RMASK = INF_RING - 1u;
struct InfOut {
    uint8_t* ring;
    uint8_t* data;          // where the member's first byte goes in global memory
    uint32_t rbase;         // ring index of data[0]
    uint32_t pos;           // bytes produced so far
    uint32_t flushed;       // data[0 .. flushed) is in global memory
    uint32_t out_len;       // the member's ISIZE
};
SVB_HD InfOut inf_out(uint8_t* ring, uint8_t* data, uint32_t out_len) {
    InfOut o;
    o.ring = ring;
    o.data = data;
    o.rbase = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(data)) & INF_RMASK;
    o.pos = 0;
    o.flushed = 0;
    o.out_len = out_len;
    return o;
}

// data[flushed .. upto) <- ring, by n_lanes lanes; the caller synchronises around it and sets o.flushed = upto afterwards
SVB_HD void inf_flush(const InfOut& o, uint32_t upto, uint32_t lane, uint32_t n_lanes) {
    const uint32_t f = o.flushed;
    const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(o.data + f)) & 15u;
    uint32_t a = mis ? f + (16u - mis) : f;                       // first 16-byte boundary of global memory
    if (a > upto) a = upto;
    for (uint32_t p = f + lane; p < a; p += n_lanes) o.data[p] = o.ring[(o.rbase + p) & INF_RMASK];
    const uint32_t n_vec = (upto - a) / 16u;
    for (uint32_t v = lane; v < n_vec; v += n_lanes) {
        const uint32_t p = a + 16u * v;
#ifdef __CUDA_ARCH__
        *reinterpret_cast<uint4*>(o.data + p) = *reinterpret_cast<const uint4*>(o.ring + ((o.rbase + p) & INF_RMASK));
#else
        for (uint32_t k = 0; k < 16u; ++k) o.data[p + k] = o.ring[(o.rbase + p + k) & INF_RMASK];
#endif
    }
    for (uint32_t p = a + 16u * n_vec + lane; p < upto; p += n_lanes) o.data[p] = o.ring[(o.rbase + p) & INF_RMASK];
}

// The lanes that run the symbol loop meet here: their shared-memory / global stores so far are visible to each other afterwards.
#ifdef __CUDA_ARCH__
#define INF_LANES_SYNC() __syncwarp()
#else
#define INF_LANES_SYNC() ((void)0)
#endif

// a match of any length, its bytes split between the n_lanes lanes; the caller synchronises the lanes before (the bytes written
// so far are the source) and after (the next symbols may read or overwrite these)
SVB_HD void inf_copy_match(const InfOut& o, uint32_t len, uint32_t dist, uint32_t lane, uint32_t n_lanes) {
    const uint32_t w = o.rbase + o.pos;
    const bool in_ring = dist + len <= INF_RING;                  // else the source has left the ring: it is in global memory, flushed
    // one round for all but the longest matches (n_lanes = 32 on the device); not unrolled: the loop is short and rare, the
    // unrolled version put twenty instructions of trip-count arithmetic in front of every match
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (uint32_t base = 0; base < len; base += n_lanes) {
        const uint32_t i = base + lane;
        if (i < len) {
            uint32_t k = i;
            if (dist < len) k = dist == 1u ? 0u : i % dist;       // an overlapping run repeats its last dist bytes
            o.ring[(w + i) & INF_RMASK] = in_ring ? o.ring[(w - dist + k) & INF_RMASK] : o.data[o.pos - dist + k];
        }
    }
}

// Symbols of the current block until its end-of-block code, run by ALL n_lanes lanes of a warp IN LOCK STEP: every lane keeps
// the same reader state and decodes the same symbol (one instruction stream: on a SIMT machine 32 lanes decoding the same
// symbol cost what one lane costs), so that every lane knows every (length, distance) without a broadcast and a match is
// copied by as many lanes as it has bytes -- six instructions instead of a byte loop of the decoding lane (measured before:
// 110 warp instructions per symbol with 1.6 active lanes, the copies of short matches most of them).  Literals are stored by
// lane 0.  When INF_FLUSH_AT bytes wait in the ring, the lanes flush them together.  On the host: lane 0 of 1.
// Returns INF_OK or an error (the same on every lane).
SVB_HD int inf_run_lanes(InfBits& b, const InfHuff& lencode, const InfHuff& distcode, const inf_len_t* tlen, const uint32_t* tdist,
                         InfOut& o, uint32_t lane, uint32_t n_lanes) {
    while (true) {
        if (o.pos - o.flushed >= INF_FLUSH_AT) {
            if (o.pos > o.out_len) return INF_ERR_OUTPUT;        // (literals are not checked one by one: nothing past out_len leaves the ring)
            INF_LANES_SYNC();
            inf_flush(o, o.pos, lane, n_lanes);
            o.flushed = o.pos;
            INF_LANES_SYNC();
        }
        if (b.cnt < 32) inf_refill(b);
        uint32_t e = tlen[static_cast<uint32_t>(b.buf) & ((1u << INF_LEN_BITS) - 1u)];
        bool fast = true;
        if ((e & 15u) == 0u) {                                   // code longer than the table
            const int sym = inf_decode(b, lencode);
            if (sym < 0) return INF_ERR_CODE;
            e = inf_len_entry(static_cast<uint32_t>(sym), 0u);
            fast = false;
        } else {
            b.buf >>= (e & 15u);
            b.cnt -= static_cast<int>(e & 15u);
        }
        if ((e & INF_L_LEN) == 0u) {                             // literal
            if (lane == 0u) o.ring[(o.rbase + o.pos) & INF_RMASK] = static_cast<uint8_t>(e >> 8);
            ++o.pos;
            // A literal from the table used at most INF_LEN_BITS of at least 32 valid bits: up to two more table literals
            // can follow without a refill, and three bytes more do not matter to the flush test (literals come in runs:
            // sequence and quality bytes).
            if (fast) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
                for (int more = 0; more < 2; ++more) {
                    const uint32_t e2 = tlen[static_cast<uint32_t>(b.buf) & ((1u << INF_LEN_BITS) - 1u)];
                    if ((e2 & 15u) == 0u || (e2 & INF_L_LEN) != 0u) break;
                    b.buf >>= (e2 & 15u);
                    b.cnt -= static_cast<int>(e2 & 15u);
                    if (lane == 0u) o.ring[(o.rbase + o.pos) & INF_RMASK] = static_cast<uint8_t>(e2 >> 8);
                    ++o.pos;
                }
            }
            continue;
        }
        const uint32_t xl = (e >> 5) & 7u;
        if (xl == 7u) {
            if (e >> 8) return INF_ERR_CODE;                     // symbol 286 / 287
            return inf_overrun(b) ? INF_ERR_INPUT : INF_OK;      // end of block
        }
        const uint32_t len = (e >> 8) + 3u + (static_cast<uint32_t>(b.buf) & ((1u << xl) - 1u));
        b.buf >>= xl;
        b.cnt -= static_cast<int>(xl);
        if (b.cnt < 32) inf_refill(b);
        uint32_t d = tdist[static_cast<uint32_t>(b.buf) & ((1u << INF_DIST_BITS) - 1u)];
        if ((d & 15u) == 0u) {
            const int dsym = inf_decode(b, distcode);
            if (dsym < 0) return INF_ERR_CODE;
            d = inf_dist_entry(static_cast<uint32_t>(dsym), 0u);
        } else {
            b.buf >>= (d & 15u);
            b.cnt -= static_cast<int>(d & 15u);
        }
        if ((d & INF_K_MASK) == INF_K_BAD) return INF_ERR_CODE;
        const uint32_t xd = (d >> 8) & 15u;
        const uint32_t dist = (d >> 16) + (static_cast<uint32_t>(b.buf) & ((1u << xd) - 1u));
        b.buf >>= xd;
        b.cnt -= static_cast<int>(xd);
        if (dist > o.pos) return INF_ERR_CODE;
        if (o.pos + len > o.out_len) return INF_ERR_OUTPUT;
        INF_LANES_SYNC();
        inf_copy_match(o, len, dist, lane, n_lanes);
        INF_LANES_SYNC();
        o.pos += len;
    }
}

// Header of the next block, read by the decoding lane: *last, *type; a stored block (type 0) reports where its bytes are
// (*st_off from `src`, *st_len) and the reader is moved past them; for Huffman blocks the canonical codes are constructed
// in lencode / distcode (`lengths` is 320 bytes of scratch).  Returns INF_OK or an error.
SVB_HD int inf_block_header(InfBits& b, const uint8_t* src, uint32_t pos, uint32_t out_len, InfHuff& lencode, InfHuff& distcode,
                            uint8_t* lengths, int* last, uint32_t* type, uint32_t* st_off, uint32_t* st_len) {
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    *last = static_cast<int>(inf_bits(b, 1));
    *type = inf_bits(b, 2);
    *st_off = 0;
    *st_len = 0;
    if (*type == 0u) {
        if (inf_overrun(b)) return INF_ERR_INPUT;
        b.p -= (b.cnt - b.virt) >> 3;                  // give back whole bytes of look-ahead, drop the rest of the current byte
        b.buf = 0;
        b.cnt = 0;
        b.virt = 0;
        if (b.end - b.p < 4) return INF_ERR_INPUT;
        const uint32_t len = b.p[0] | (static_cast<uint32_t>(b.p[1]) << 8);
        const uint32_t nlen = b.p[2] | (static_cast<uint32_t>(b.p[3]) << 8);
        b.p += 4;
        if (len != (~nlen & 0xFFFFu)) return INF_ERR_CODE;
        if (static_cast<uint32_t>(b.end - b.p) < len) return INF_ERR_INPUT;
        if (pos + len > out_len) return INF_ERR_OUTPUT;
        *st_off = static_cast<uint32_t>(b.p - src);
        *st_len = len;
        b.p += len;
        return INF_OK;
    }
    if (*type == 1u) {
        int s = 0;
        for (; s < 144; ++s) lengths[s] = 8;
        for (; s < 256; ++s) lengths[s] = 9;
        for (; s < 280; ++s) lengths[s] = 7;
        for (; s < 288; ++s) lengths[s] = 8;
        inf_construct(lencode, lengths, 288);
        for (s = 0; s < 30; ++s) lengths[s] = 5;
        inf_construct(distcode, lengths, 30);
        return INF_OK;
    }
    if (*type != 2u) return INF_ERR_CODE;
    const int nlen = static_cast<int>(inf_bits(b, 5)) + 257;
    const int ndist = static_cast<int>(inf_bits(b, 5)) + 1;
    const int ncode = static_cast<int>(inf_bits(b, 4)) + 4;
    if (nlen > 286 || ndist > 30) return INF_ERR_CODE;
    int idx = 0;
    for (; idx < ncode; ++idx) lengths[order[idx]] = static_cast<uint8_t>(inf_bits(b, 3));
    for (; idx < 19; ++idx) lengths[order[idx]] = 0;
    if (inf_construct(lencode, lengths, 19) != 0) return INF_ERR_CODE;      // the code-length code must be complete
    idx = 0;
    while (idx < nlen + ndist) {
        const int sym = inf_decode(b, lencode);
        if (sym < 0) return INF_ERR_CODE;
        if (sym < 16) {
            lengths[idx++] = static_cast<uint8_t>(sym);
        } else {
            int len = 0, rep;
            if (sym == 16) {
                if (idx == 0) return INF_ERR_CODE;
                len = lengths[idx - 1];
                rep = 3 + static_cast<int>(inf_bits(b, 2));
            } else if (sym == 17) {
                rep = 3 + static_cast<int>(inf_bits(b, 3));
            } else {
                rep = 11 + static_cast<int>(inf_bits(b, 7));
            }
            if (idx + rep > nlen + ndist) return INF_ERR_CODE;
            while (rep--) lengths[idx++] = static_cast<uint8_t>(len);
        }
    }
    if (lengths[256] == 0) return INF_ERR_CODE;         // no end-of-block code
    // the distance code first: constructing the length code afterwards does not disturb `lengths + nlen`
    int e1 = inf_construct(distcode, lengths + nlen, ndist);
    if (e1 < 0 || (e1 > 0 && ndist - distcode.count[0] != 1)) return INF_ERR_CODE;
    e1 = inf_construct(lencode, lengths, nlen);
    if (e1 < 0 || (e1 > 0 && nlen - lencode.count[0] != 1)) return INF_ERR_CODE;
    return INF_OK;
}
