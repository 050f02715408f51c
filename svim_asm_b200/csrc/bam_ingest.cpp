// Host ingest: BGZF/BAM file -> the flat, 16-byte aligned record image that svb_load_records uploads.
//
// Stands in for what the reference obtains from pysam/htslib (absent from this image; semantics
// from the SAM/BAM specification v1):
//   pysam.AlignmentFile(path), .header["HD"]["SO"], .references, .lengths   (svim-asm:63-67,172-173)
//   bam.fetch(contig) in header order                                        (SVIM_COLLECT.py:62-65)
//   AlignedSegment.cigartuples incl. the CG:B,I long-CIGAR convention        (SVIM_intra.py:37)
//   retrieve_other_alignments: SA:Z text -> pseudo alignments                (SVIM_COLLECT.py:8-58)
// BGZF members are independent deflate streams, so they are inflated by a pool of threads.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/svimasm_b200.h"

#include "bam_host.h"

namespace {

void set_err(char* err, int err_len, const std::string& msg) {
    if (err && err_len > 0) snprintf(err, static_cast<size_t>(err_len), "%s", msg.c_str());
}

inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t rdi32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

bool inflate_block(const uint8_t* src, uint64_t src_len, uint8_t* dst, uint64_t dst_len) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<Bytef*>(src);
    zs.avail_in = static_cast<uInt>(src_len);
    zs.next_out = dst;
    zs.avail_out = static_cast<uInt>(dst_len);
    const int rc = inflate(&zs, Z_FINISH);
    const bool ok = rc == Z_STREAM_END && zs.total_out == dst_len;
    inflateEnd(&zs);
    return ok;
}

// python int(): optional surrounding blanks, optional sign, decimal digits
bool parse_pyint(const char* b, const char* e, long long* out) {
    while (b < e && (*b == ' ' || *b == '\t')) ++b;
    while (e > b && (e[-1] == ' ' || e[-1] == '\t')) --e;
    if (b == e) return false;
    bool neg = false;
    if (*b == '+' || *b == '-') { neg = *b == '-'; ++b; }
    if (b == e) return false;
    long long v = 0;
    for (; b < e; ++b) {
        if (*b < '0' || *b > '9') return false;
        if (v < (1ll << 58)) v = v * 10 + (*b - '0');
    }
    *out = neg ? -v : v;
    return true;
}

int op_code(char c) {
    static const char letters[] = "MIDNSHP=XB";
    const char* p = strchr(letters, c);
    return (p && c) ? static_cast<int>(p - letters) : -1;
}

// One SA element "rname,pos,strand,CIGAR,mapQ,NM" -> segment.  1 = appended, 0 = skipped, <0 = error.
int parse_sa_element(const char* b, const char* e, const char* const* names, int32_t n_contig, svb_segment* out) {
    const char* f[8];
    int nf = 0;
    f[nf++] = b;
    for (const char* p = b; p < e; ++p)
        if (*p == ',') {
            if (nf < 8) f[nf] = p + 1;
            ++nf;
        }
    if (nf != 6) return 0;                                              // SVIM_COLLECT.py:22-23
    auto fend = [&](int i) { return i + 1 < 6 ? f[i + 1] - 1 : e; };
    long long pos = 0, mapq = 0, nm = 0;
    if (!parse_pyint(f[1], fend(1), &pos) || !parse_pyint(f[4], fend(4), &mapq) || !parse_pyint(f[5], fend(5), &nm))
        return SVB_ERR_FORMAT;                                          // int() raises ValueError: the reference aborts
    svb_segment s;
    memset(&s, 0, sizeof s);
    s.tid = -1;
    const size_t name_len = static_cast<size_t>(fend(0) - f[0]);
    for (int32_t t = 0; t < n_contig; ++t)
        if (strlen(names[t]) == name_len && memcmp(names[t], f[0], name_len) == 0) { s.tid = t; break; }
    s.pos = static_cast<int32_t>(pos - 1);                              // :41
    s.is_reverse = !(fend(2) - f[2] == 1 && *f[2] == '+');              // :36-39
    s.mapq = (mapq < 0 || mapq > 255) ? 0 : static_cast<uint8_t>(mapq);  // :42-45
    // CIGAR text: pysam keeps every "<digits><op>" it finds (regex findall), ignoring anything else
    long long ref_span = 0, read_len = 0, qas = 0, qae = 0;
    bool lead = true, any = false;
    long long num = 0;
    bool have = false;
    for (const char* p = f[3]; p < fend(3); ++p) {
        if (*p >= '0' && *p <= '9') {
            if (num < (1ll << 40)) num = num * 10 + (*p - '0');
            have = true;
            continue;
        }
        const int op = op_code(*p);
        if (op >= 0 && have) {
            if (num >= (1ll << 28)) return 0;                           // OverflowError -> logged and skipped (:48-50)
            any = true;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_span += num;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8 || op == 5) read_len += num;
            if (lead) {
                if (op == 4) qas += num;
                else if (op != 5) lead = false;
            }
            if (op == 0 || op == 1 || op == 7 || op == 8 || (op == 4 && qae == 0)) qae += num;
        }
        num = 0;
        have = false;
    }
    if (!any) { ref_span = 0; }                                         // empty CIGAR: reference_end is None; never valid input
    s.ref_end = static_cast<int32_t>(s.pos + (ref_span > 0 ? ref_span : 1));
    s.q_astart = static_cast<int32_t>(qas);
    s.q_aend = static_cast<int32_t>(qae);
    s.read_len = static_cast<int32_t>(read_len);
    *out = s;
    return 1;
}

}  // namespace

namespace {

// BSIZE + 1 of the BGZF member whose gzip header starts at raw[in], 0 when there is no well-formed member there
inline uint64_t bgzf_member_size(const uint8_t* raw, uint64_t size, uint64_t in, uint64_t* hdr_len) {
    if (in + 18 > size) return 0;
    const uint8_t* h = raw + in;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
    const uint16_t xlen = rd16(h + 10);
    if (in + 12ull + xlen > size) return 0;
    uint64_t bsize = 0;
    for (uint32_t x = 0; x + 4 <= xlen;) {
        const uint8_t* sub = h + 12 + x;
        const uint16_t slen = rd16(sub + 2);
        if (sub[0] == 'B' && sub[1] == 'C' && slen == 2) bsize = static_cast<uint64_t>(rd16(sub + 4)) + 1;
        x += 4u + slen;
    }
    *hdr_len = 12ull + xlen;
    if (!bsize || bsize < *hdr_len + 8 || in + bsize > size) return 0;
    return bsize;
}

// the members that START in [from, upto): appended to `out` (out_off left 0); returns where the chain goes on (>= upto, or
// == size at the end of the file), or UINT64_MAX when a member is malformed
uint64_t bgzf_walk(const uint8_t* raw, uint64_t size, uint64_t from, uint64_t upto, std::vector<BgzfMember>* out) {
    uint64_t in = from;
    while (in < upto && in + 18 <= size) {
        uint64_t hdr_len = 0;
        const uint64_t bsize = bgzf_member_size(raw, size, in, &hdr_len);
        if (!bsize) return UINT64_MAX;
        BgzfMember b;
        b.in_off = in + hdr_len;
        b.file_off = in;
        b.in_len = bsize - hdr_len - 8;
        b.out_len = rd32(raw + in + bsize - 4);
        b.out_off = 0;
        out->push_back(b);
        in += bsize;
    }
    return in;
}

}  // namespace

// The member table is a linked list through the file (every header holds the size of its member): 43,000 hops for a
// whole-genome BAM, each a page fault of the mapping when the file is only in the page cache -- 20 ms for one thread.
// With n_threads > 1 the file is cut into ranges; every thread but the first looks for the first offset of its range
// where a chain of three well-formed members starts and walks on from there; the pieces are accepted when each range's
// chain ends exactly where the next one began (a false start inside compressed data does not survive that), else the
// table is built serially.
bool bgzf_member_table(const uint8_t* raw, uint64_t size, std::vector<BgzfMember>* blocks, uint64_t* total_out_p, std::string* why, int n_threads) {
    std::vector<BgzfMember> all;
    bool have = false;
    const int T = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(std::max(n_threads, 1)), size >> 23));      // ranges of at least 8 MB
    if (T > 1) {
        std::vector<std::vector<BgzfMember>> part(static_cast<size_t>(T));
        std::vector<uint64_t> first(static_cast<size_t>(T), UINT64_MAX), next(static_cast<size_t>(T), UINT64_MAX);
        auto work = [&](int i) {
            const uint64_t lo = size * static_cast<uint64_t>(i) / T, hi = size * static_cast<uint64_t>(i + 1) / T;
            uint64_t start = lo;
            if (i > 0) {
                start = UINT64_MAX;
                for (uint64_t at = lo; at < hi && at + 18 <= size; ++at) {
                    if (raw[at] != 0x1f || raw[at + 1] != 0x8b) continue;
                    uint64_t x = at, hl = 0;
                    int good = 0;
                    for (; good < 3 && x < size; ++good) {
                        const uint64_t bs = bgzf_member_size(raw, size, x, &hl);
                        if (!bs) break;
                        x += bs;
                    }
                    if (good == 3 || (good > 0 && x == size)) { start = at; break; }
                }
                if (start == UINT64_MAX) return;
            }
            first[static_cast<size_t>(i)] = start;
            next[static_cast<size_t>(i)] = bgzf_walk(raw, size, start, i + 1 == T ? size : hi, &part[static_cast<size_t>(i)]);
        };
        std::vector<std::thread> pool;
        for (int i = 1; i < T; ++i) pool.emplace_back(work, i);
        work(0);
        for (auto& th : pool) th.join();
        have = true;
        for (int i = 0; i < T && have; ++i) {
            if (first[static_cast<size_t>(i)] == UINT64_MAX || next[static_cast<size_t>(i)] == UINT64_MAX) have = false;
            else if (i + 1 < T && next[static_cast<size_t>(i)] != first[static_cast<size_t>(i + 1)]) have = false;
        }
        if (have && (next[static_cast<size_t>(T - 1)] + 18 <= size)) have = false;      // (cannot happen: the last range walks to the end)
        if (have)
            for (auto& p : part) all.insert(all.end(), p.begin(), p.end());
    }
    if (!have) {
        all.clear();
        // the serial walk; it also words the errors
        uint64_t in = 0;
        while (in + 18 <= size) {
            const uint8_t* h = raw + in;
            if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { *why = "not a BGZF file"; return false; }
            uint64_t hdr_len = 0;
            const uint64_t bsize = bgzf_member_size(raw, size, in, &hdr_len);
            if (!bsize) { *why = "truncated BGZF block"; return false; }
            if (bgzf_walk(raw, size, in, in + 1, &all) == UINT64_MAX) { *why = "truncated BGZF block"; return false; }
            in += bsize;
        }
    }
    uint64_t total_out = 0;
    blocks->reserve(blocks->size() + all.size());
    for (BgzfMember& b : all) {
        b.out_off = total_out;
        total_out += b.out_len;
        if (b.out_len) blocks->push_back(b);
    }
    *total_out_p = total_out;
    return true;
}

int64_t bam_parse_header(const uint8_t* p0, uint64_t avail, svb_bam* bam, std::string* why) {
    const uint8_t* p = p0;
    const uint8_t* const end = p0 + avail;
    if (avail < 12 || memcmp(p, "BAM\1", 4) != 0) { *why = "not a BAM file"; return -1; }
    const int32_t l_text = rdi32(p + 4);
    if (l_text < 0) { *why = "bad BAM header"; return -1; }
    if (p + 12 + l_text > end) { *why = "bad BAM header"; return -2; }       // -2: the caller may retry with more bytes
    {
        const std::string text(reinterpret_cast<const char*>(p + 8), strnlen(reinterpret_cast<const char*>(p + 8), static_cast<size_t>(l_text)));
        size_t ls = 0;
        while (ls < text.size()) {
            size_t le = text.find('\n', ls);
            if (le == std::string::npos) le = text.size();
            if (text.compare(ls, 3, "@HD") == 0) {
                size_t so = text.find("\tSO:", ls);
                if (so != std::string::npos && so < le) {
                    size_t ve = text.find_first_of("\t\n\r", so + 4);
                    if (ve == std::string::npos || ve > le) ve = le;
                    bam->sort_order = text.substr(so + 4, ve - so - 4);
                }
            }
            ls = le + 1;
        }
    }
    p += 8 + l_text;
    const int32_t n_ref = rdi32(p);
    p += 4;
    bam->contig_names.clear();
    bam->contig_len.clear();
    for (int32_t i = 0; i < n_ref; ++i) {
        if (p + 4 > end) { *why = "bad BAM reference list"; return -2; }
        const int32_t l_name = rdi32(p);
        if (l_name < 1) { *why = "bad BAM reference list"; return -1; }
        if (p + 8 + l_name > end) { *why = "bad BAM reference list"; return -2; }
        bam->contig_names.emplace_back(reinterpret_cast<const char*>(p + 4), static_cast<size_t>(l_name - 1));
        bam->contig_len.push_back(rdi32(p + 4 + l_name));
        p += 8 + l_name;
    }
    return p - p0;
}

extern "C" {

int svb_parse_sa(const char* sa_text, const char* const* contig_names, int32_t n_contig, svb_segment* out, int32_t cap) {
    if (!sa_text || (n_contig && !contig_names)) return SVB_ERR_ARG;
    int n = 0;
    const char* b = sa_text;
    const char* end = sa_text + strlen(sa_text);
    while (b <= end) {
        const char* e = static_cast<const char*>(memchr(b, ';', static_cast<size_t>(end - b)));
        if (!e) e = end;
        svb_segment s;
        const int rc = parse_sa_element(b, e, contig_names, n_contig, &s);
        if (rc < 0) return rc;
        if (rc == 1) {
            if (n < cap && out) out[n] = s;
            ++n;
        }
        if (e == end) break;
        b = e + 1;
    }
    return n;
}

int svb_bam_open(const char* path, int n_threads, svb_bam** out, char* err, int err_len) {
    if (!path || !out) return SVB_ERR_ARG;
    *out = nullptr;
    FILE* fp = fopen(path, "rb");
    if (!fp) { set_err(err, err_len, std::string("cannot open ") + path); return SVB_ERR_IO; }
    fseek(fp, 0, SEEK_END);
    const long fsize = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<uint8_t> raw(static_cast<size_t>(fsize));
    if (fsize && fread(raw.data(), 1, raw.size(), fp) != raw.size()) { fclose(fp); set_err(err, err_len, "short read"); return SVB_ERR_IO; }
    fclose(fp);

    std::vector<BgzfMember> blocks;
    uint64_t total_out = 0;
    {
        std::string why;
        if (!bgzf_member_table(raw.data(), raw.size(), &blocks, &total_out, &why)) { set_err(err, err_len, why); return SVB_ERR_IO; }
    }
    std::vector<uint8_t> data(total_out);
    {
        std::atomic<size_t> next(0);
        std::atomic<bool> bad(false);
        const int nt = n_threads > 0 ? n_threads : static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
        auto work = [&]() {
            for (size_t i = next.fetch_add(1); i < blocks.size(); i = next.fetch_add(1)) {
                const BgzfMember& b = blocks[i];
                if (!inflate_block(raw.data() + b.in_off, b.in_len, data.data() + b.out_off, b.out_len)) bad = true;
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < nt && static_cast<size_t>(t) < blocks.size(); ++t) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
        if (bad) { set_err(err, err_len, "inflate failed (corrupt BGZF block)"); return SVB_ERR_IO; }
    }
    std::vector<uint8_t>().swap(raw);

    // ---- BAM header
    svb_bam* bam = new (std::nothrow) svb_bam();
    if (!bam) return SVB_ERR_NOMEM;
    const uint8_t* p = data.data();
    const uint8_t* const end = p + data.size();
    {
        std::string why;
        const int64_t used = bam_parse_header(p, data.size(), bam, &why);
        if (used < 0) { delete bam; set_err(err, err_len, why); return SVB_ERR_IO; }
        p += used;
    }
    const int32_t n_ref = static_cast<int32_t>(bam->contig_names.size());
    std::vector<const char*> name_ptrs;
    for (auto& s : bam->contig_names) name_ptrs.push_back(s.c_str());

    // ---- records: first pass finds boundaries and sizes
    struct Rec { const uint8_t* body; uint32_t len; const uint8_t* cg; uint32_t cg_n; const char* sa; };
    std::vector<Rec> recs;
    uint64_t ops_padded = 0, seq_bytes = 0, name_bytes = 0;
    while (p + 4 <= end) {
        const int32_t block = rdi32(p);
        if (block < 32 || p + 4 + block > end) { delete bam; set_err(err, err_len, "truncated BAM record"); return SVB_ERR_IO; }
        const uint8_t* body = p + 4;
        const uint8_t l_name = body[8];
        const uint16_t n_cig = rd16(body + 12);
        const int32_t l_seq = rdi32(body + 16);
        const uint8_t* cig = body + 32 + l_name;
        const uint8_t* tags = cig + 4ull * n_cig + (static_cast<uint64_t>(l_seq) + 1) / 2 + static_cast<uint64_t>(l_seq);
        if (l_seq < 0 || tags > body + block) { delete bam; set_err(err, err_len, "malformed BAM record"); return SVB_ERR_IO; }
        Rec r{body, static_cast<uint32_t>(block), nullptr, 0, nullptr};
        const uint8_t* t = tags;
        const uint8_t* te = body + block;
        while (t + 3 <= te) {
            const char a0 = static_cast<char>(t[0]), a1 = static_cast<char>(t[1]), ty = static_cast<char>(t[2]);
            t += 3;
            size_t adv = 0;
            switch (ty) {
                case 'A': case 'c': case 'C': adv = 1; break;
                case 's': case 'S': adv = 2; break;
                case 'i': case 'I': case 'f': adv = 4; break;
                case 'Z': case 'H': {
                    const uint8_t* z = static_cast<const uint8_t*>(memchr(t, 0, static_cast<size_t>(te - t)));
                    if (!z) { delete bam; set_err(err, err_len, "unterminated tag"); return SVB_ERR_IO; }
                    if (a0 == 'S' && a1 == 'A' && ty == 'Z') r.sa = reinterpret_cast<const char*>(t);
                    adv = static_cast<size_t>(z - t) + 1;
                    break;
                }
                case 'B': {
                    if (t + 5 > te) { delete bam; set_err(err, err_len, "bad B tag"); return SVB_ERR_IO; }
                    const char sub = static_cast<char>(t[0]);
                    const uint32_t cnt = rd32(t + 1);
                    const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                    if (a0 == 'C' && a1 == 'G' && sub == 'I') { r.cg = t + 5; r.cg_n = cnt; }
                    adv = 5 + es * cnt;
                    break;
                }
                default: delete bam; set_err(err, err_len, "unknown tag type"); return SVB_ERR_IO;
            }
            if (t + adv > te) { delete bam; set_err(err, err_len, "tag overruns record"); return SVB_ERR_IO; }
            t += adv;
        }
        // long CIGAR convention: "<l_seq>S<ref_len>N" placeholder + CG:B,I
        uint32_t real_n = n_cig;
        if (r.cg && n_cig == 2 && rd32(cig) == ((static_cast<uint32_t>(l_seq) << 4) | 4u) && (rd32(cig + 4) & 15u) == 3u) real_n = r.cg_n;
        else r.cg = nullptr;
        ops_padded += (static_cast<uint64_t>(real_n) + 3) / 4 * 4;
        seq_bytes += (static_cast<uint64_t>(l_seq) + 1) / 2;
        name_bytes += l_name;
        recs.push_back(r);
        p += 4 + block;
    }

    // ---- second pass: fill the flat arrays
    const size_t n = recs.size();
    bam->hdr.resize(n);
    bam->sa_count.assign(n, 0);
    bam->cigar.assign(ops_padded, 15u);
    bam->seq4.resize(seq_bytes);
    bam->seq_off.resize(n + 1);
    bam->names.resize(name_bytes);
    bam->name_off.resize(n + 1);
    bam->sa_text_off.assign(n, -1);
    uint64_t co = 0, so = 0, no = 0;
    for (size_t i = 0; i < n; ++i) {
        const Rec& r = recs[i];
        const uint8_t* body = r.body;
        const uint8_t l_name = body[8];
        const uint16_t n_cig = rd16(body + 12);
        const int32_t l_seq = rdi32(body + 16);
        const uint8_t* cig = body + 32 + l_name;
        svb_aln_hdr& h = bam->hdr[i];
        memset(&h, 0, sizeof h);
        h.tid = rdi32(body);
        h.pos = rdi32(body + 4);
        h.mapq = body[9];
        h.flag = rd16(body + 14);
        h.l_seq = static_cast<uint32_t>(l_seq);
        h.cigar_off = co;
        if (r.cg) {
            h.n_cigar = r.cg_n;
            memcpy(bam->cigar.data() + co, r.cg, 4ull * r.cg_n);
        } else {
            h.n_cigar = n_cig;
            memcpy(bam->cigar.data() + co, cig, 4ull * n_cig);
        }
        co += (static_cast<uint64_t>(h.n_cigar) + 3) / 4 * 4;
        bam->seq_off[i] = so;
        memcpy(bam->seq4.data() + so, cig + 4ull * n_cig, (static_cast<uint64_t>(l_seq) + 1) / 2);
        so += (static_cast<uint64_t>(l_seq) + 1) / 2;
        bam->name_off[i] = no;
        memcpy(bam->names.data() + no, body + 32, l_name);
        no += l_name;
        h.sa_first = static_cast<uint32_t>(bam->seg.size());
        if (r.sa) {
            bam->sa_text_off[i] = static_cast<int64_t>(bam->sa_text.size());
            bam->sa_text.insert(bam->sa_text.end(), r.sa, r.sa + strlen(r.sa) + 1);
            const int cnt = svb_parse_sa(r.sa, name_ptrs.data(), n_ref, nullptr, 0);
            if (cnt < 0) {
                delete bam;
                set_err(err, err_len, std::string("malformed SA tag (int() would raise): ") + r.sa);
                return cnt;
            }
            bam->seg.resize(bam->seg.size() + static_cast<size_t>(cnt));
            svb_parse_sa(r.sa, name_ptrs.data(), n_ref, bam->seg.data() + h.sa_first, cnt);
            bam->sa_count[i] = static_cast<uint32_t>(cnt);
        }
    }
    bam->seq_off[n] = so;
    bam->name_off[n] = no;
    *out = bam;
    return SVB_OK;
}

void svb_bam_close(svb_bam* bam) { delete bam; }
int64_t svb_bam_n_records(const svb_bam* b) { return b ? static_cast<int64_t>(b->hdr.size()) : -1; }
int64_t svb_bam_n_ops_padded(const svb_bam* b) { return b ? static_cast<int64_t>(b->cigar.size()) : -1; }
int64_t svb_bam_n_segments(const svb_bam* b) { return b ? static_cast<int64_t>(b->seg.size()) : -1; }
int32_t svb_bam_n_contigs(const svb_bam* b) { return b ? static_cast<int32_t>(b->contig_names.size()) : -1; }
const char* svb_bam_contig_name(const svb_bam* b, int32_t tid) {
    return (b && tid >= 0 && static_cast<size_t>(tid) < b->contig_names.size()) ? b->contig_names[static_cast<size_t>(tid)].c_str() : nullptr;
}
const int32_t* svb_bam_contig_lengths(const svb_bam* b) { return b ? b->contig_len.data() : nullptr; }
const char* svb_bam_sort_order(const svb_bam* b) { return b ? b->sort_order.c_str() : ""; }
const svb_aln_hdr* svb_bam_headers(const svb_bam* b) { return b ? b->hdr.data() : nullptr; }
const uint32_t* svb_bam_cigar(const svb_bam* b) { return b ? b->cigar.data() : nullptr; }
const svb_segment* svb_bam_segments(const svb_bam* b) { return b ? b->seg.data() : nullptr; }
const uint32_t* svb_bam_sa_count(const svb_bam* b) { return b ? b->sa_count.data() : nullptr; }
const uint8_t* svb_bam_seq4(const svb_bam* b) { return b ? b->seq4.data() : nullptr; }
const uint64_t* svb_bam_seq_offsets(const svb_bam* b) { return b ? b->seq_off.data() : nullptr; }
const char* svb_bam_sa_text(const svb_bam* b, int64_t record) {
    if (!b || record < 0 || static_cast<size_t>(record) >= b->hdr.size() || b->sa_text_off[static_cast<size_t>(record)] < 0) return nullptr;
    return b->sa_text.data() + b->sa_text_off[static_cast<size_t>(record)];
}
const char* svb_bam_query_name(const svb_bam* b, int64_t record) {
    return (b && record >= 0 && static_cast<size_t>(record) < b->hdr.size()) ? b->names.data() + b->name_off[static_cast<size_t>(record)] : nullptr;
}

// test hook (svimasm_b200_debug.h): the BGZF member table of a file built with n_threads; out = {members, inflated bytes, a
// checksum over every member's offsets and sizes}
int svb_bgzf_member_table_check(const char* path, int n_threads, uint64_t out[3]) {
    if (!path || !out) return SVB_ERR_ARG;
    FILE* fp = fopen(path, "rb");
    if (!fp) return SVB_ERR_IO;
    fseek(fp, 0, SEEK_END);
    const long fsize = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<uint8_t> raw(static_cast<size_t>(std::max(fsize, 0l)));
    const bool ok = !fsize || fread(raw.data(), 1, raw.size(), fp) == raw.size();
    fclose(fp);
    if (!ok) return SVB_ERR_IO;
    std::vector<BgzfMember> blocks;
    uint64_t total_out = 0;
    std::string why;
    if (!bgzf_member_table(raw.data(), raw.size(), &blocks, &total_out, &why, n_threads)) return SVB_ERR_IO;
    uint64_t sum = 1469598103934665603ull;
    for (const BgzfMember& b : blocks)
        for (uint64_t v : {b.in_off, b.in_len, b.out_off, b.out_len, b.file_off}) sum = (sum ^ v) * 1099511628211ull;
    out[0] = blocks.size();
    out[1] = total_out;
    out[2] = sum;
    return SVB_OK;
}

}  // extern "C"
