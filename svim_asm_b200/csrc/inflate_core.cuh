// DEFLATE (RFC 1951) decoder for ONE BGZF member, written to run as one GPU thread per member.
//
// Where it sits: the reference opens its input with pysam.AlignmentFile (svim-asm:63,85-86), i.e. htslib inflates
// the BGZF blocks of the BAM file on the host [ext].  BGZF members are independent raw-deflate streams of at most
// 64 KiB each, so a whole-genome BAM is tens of thousands of independent jobs: one thread per member, every member
// in flight at once.  The decoder keeps its state small (canonical-Huffman count/symbol arrays, 0.7 KB) so that a
// few hundred threads per SM keep theirs in L1; it reads the compressed stream and writes the output bytewise.
//
// The same code compiles for the host (tests/hostcheck) where it is checked against zlib.
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
