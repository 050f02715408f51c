// VCF body lines of the candidate table, as a PLAN of byte pieces per record.
//
// Where it sits: the step after the hot path.  write_final_vcf (reference SVIM_COMBINE.py:428-475) asks every
// candidate for its text (get_vcf_entry*, SVCandidate.py:53-78,99-125,151-176,203-261,296-347,389-443); with sequence
// alleles that is 4-6 reference.fetch calls and a query_sequence slice per record, and the REF / ALT strings are
// most of the file (a deleted region, an inserted sequence, an inverted region).  Here one record is a short list
// of pieces -- literal text, a contig name, a run of the HBM-resident reference (plain, reverse-complemented or
// repeated), a run of a 4-bit query sequence -- so its length is known without touching any bases, and its bytes
// can be produced independently of every other record: one warp per record on the device (vcf_device.cu).
//
// The plan and the per-byte getter compile for the host as well (tests/hostcheck), where they are compared with the
// python writer line by line.
#pragma once
#include <stdint.h>

#include "../../include/svimasm_b200.h"
#include "linkage.cuh"   // SVB_HD

// how a table row is written; chosen on the host from the row type and the writer's options
enum : uint32_t {
    VCF_DEL = 0, VCF_INV = 1, VCF_INS = 2,
    VCF_TAN_AS_INS = 3, VCF_TAN_AS_DUP = 4,         // get_vcf_entry_as_ins / _as_dup of CandidateDuplicationTandem
    VCF_INT_AS_INS = 5, VCF_INT_AS_DUP = 6,         // ... of CandidateDuplicationInterspersed
    VCF_BND = 7, VCF_BND_MATE = 8,                  // get_vcf_entry / get_vcf_entry_reverse of CandidateBreakend
    VCF_MODE_COUNT = 9
};
enum : uint32_t { VCF_SYMBOLIC = 1u };              // options.symbolic_alleles: REF "N", ALT "<TYPE>"

typedef svb_vcf_entry VcfEntry;            // {row, id (1-based running number of its ID label), mode}
static_assert(int(VCF_BND_MATE) == int(SVB_VCF_BND_MATE) && int(VCF_TAN_AS_DUP) == int(SVB_VCF_TAN_AS_DUP) && int(VCF_SYMBOLIC) == int(SVB_VCF_SYMBOLIC), "public enum");

enum : uint8_t { VP_TEXT = 0, VP_NAME = 1, VP_REF = 2, VP_REF_REVCOMP = 3, VP_SEQ4 = 4 };
struct VcfPiece {
    uint64_t src;        // TEXT: offset in plan.text; NAME: offset in the name blob; REF*: base index; SEQ4: base index in the record
    uint64_t len;        // bytes this piece contributes
    uint32_t unit;       // REF: period (the run repeats when len > unit)
    uint32_t aux;        // SEQ4: record index
    uint8_t kind;
    uint8_t slot;        // SEQ4: haplotype of the record image (0 haploid, 1, 2)
};
constexpr int VCF_MAX_PIECES = 8;
constexpr int VCF_TEXT_CAP = 240;
struct VcfPlan {
    VcfPiece piece[VCF_MAX_PIECES];
    uint32_t n_pieces;
    uint32_t n_text;
    char text[VCF_TEXT_CAP];
};

struct VcfEnv {
    const uint8_t* bases;            // upper-cased reference, contigs concatenated
    const uint64_t* contig_off;      // [n_contig + 1]
    int32_t n_contig;
    const uint8_t* names;            // contig names, concatenated
    const uint32_t* name_off;        // [n_contig + 1]
    const uint8_t* seq4[3];          // 4-bit query sequences by haplotype slot (null when absent)
    const uint64_t* seq_off[3];      // [n_aln + 1] byte offsets
    uint32_t flags;                  // VCF_SYMBOLIC
};

// ---- plan construction ------------------------------------------------------------------------------------------
SVB_HD void vcf_txt(VcfPlan& p, const char* s) {
    while (*s && p.n_text < static_cast<uint32_t>(VCF_TEXT_CAP)) p.text[p.n_text++] = *s++;
}
SVB_HD void vcf_int(VcfPlan& p, int64_t v) {                     // "%d"
    char tmp[24];
    int n = 0;
    uint64_t u = v < 0 ? 0ull - static_cast<uint64_t>(v) : static_cast<uint64_t>(v);
    do {
        tmp[n++] = static_cast<char>('0' + u % 10ull);
        u /= 10ull;
    } while (u);
    if (v < 0) tmp[n++] = '-';
    while (n && p.n_text < static_cast<uint32_t>(VCF_TEXT_CAP)) p.text[p.n_text++] = tmp[--n];
}
SVB_HD void vcf_push(VcfPlan& p, uint8_t kind, uint64_t src, uint64_t len, uint32_t unit = 0, uint32_t aux = 0, uint8_t slot = 0) {
    if (len == 0 || p.n_pieces >= static_cast<uint32_t>(VCF_MAX_PIECES)) return;
    VcfPiece& q = p.piece[p.n_pieces++];
    q.src = src;
    q.len = len;
    q.unit = unit;
    q.aux = aux;
    q.kind = kind;
    q.slot = slot;
}
// literal text written since `from` becomes one piece
SVB_HD void vcf_flush(VcfPlan& p, uint32_t& from) {
    vcf_push(p, VP_TEXT, from, p.n_text - from);
    from = p.n_text;
}
SVB_HD void vcf_name(VcfPlan& p, const VcfEnv& e, int32_t tid) {
    if (tid < 0 || tid >= e.n_contig) return;
    vcf_push(p, VP_NAME, e.name_off[tid], e.name_off[tid + 1] - e.name_off[tid]);
}
// reference.fetch(contig, start, end).upper(): the run is cut at the contig's ends like a python slice
SVB_HD void vcf_ref(VcfPlan& p, const VcfEnv& e, int32_t tid, int64_t start, int64_t end, uint32_t times = 1, bool revcomp = false) {
    if (tid < 0 || tid >= e.n_contig) return;
    const int64_t len = static_cast<int64_t>(e.contig_off[tid + 1] - e.contig_off[tid]);
    if (start < 0) start = 0;
    if (end > len) end = len;
    if (end <= start || times == 0) return;
    const uint64_t unit = static_cast<uint64_t>(end - start);
    vcf_push(p, revcomp ? VP_REF_REVCOMP : VP_REF, e.contig_off[tid] + static_cast<uint64_t>(start), unit * times, static_cast<uint32_t>(unit));
}

SVB_HD const char* vcf_label(uint32_t mode) {
    switch (mode) {
        case VCF_DEL: return "DEL";
        case VCF_INV: return "INV";
        case VCF_TAN_AS_DUP: return "DUP_TANDEM";
        case VCF_INT_AS_DUP: return "DUP_INT";
        case VCF_BND: case VCF_BND_MATE: return "BND";
        default: return "INS";
    }
}
SVB_HD const char* vcf_genotype(uint32_t g) { return g == SVB_GT_HOM ? "1/1" : (g == SVB_GT_HAP1 ? "1/0" : "0/1"); }

// CHROM \t POS \t ID \t  -- the name is its own piece, the rest is text
SVB_HD void vcf_head(VcfPlan& p, const VcfEnv& e, int32_t tid, int64_t pos, const VcfEntry& en, uint32_t& from) {
    vcf_name(p, e, tid);
    vcf_txt(p, "\t");
    vcf_int(p, pos);
    vcf_txt(p, "\tsvim_asm.");
    vcf_txt(p, vcf_label(en.mode));
    vcf_txt(p, ".");
    vcf_int(p, en.id);
    vcf_txt(p, "\t");
}
// \t . \t FILTER \t INFO \t FORMAT \t SAMPLE \n
SVB_HD void vcf_tail_open(VcfPlan& p, const char* filter) {
    vcf_txt(p, "\t.\t");
    vcf_txt(p, filter);
    vcf_txt(p, "\tSVTYPE=");
}
SVB_HD void vcf_end_svlen(VcfPlan& p, int64_t end, int64_t svlen) {
    vcf_txt(p, "END=");
    vcf_int(p, end);
    vcf_txt(p, ";SVLEN=");
    vcf_int(p, svlen);
}
SVB_HD void vcf_tail_close(VcfPlan& p, const svb_row& r, uint32_t& from) {
    vcf_txt(p, "\tGT\t");
    vcf_txt(p, vcf_genotype(r.genotype));
    vcf_txt(p, "\n");
    vcf_flush(p, from);
}

SVB_HD void vcf_plan(const svb_row& r, const VcfEntry& en, const VcfEnv& e, VcfPlan& p) {
    p.n_pieces = 0;
    p.n_text = 0;
    uint32_t from = 0;
    const bool alleles = (e.flags & VCF_SYMBOLIC) == 0u;
    const int64_t s = r.src_start, t = r.src_end, ds = r.dst_start, dt = r.dst_end;
    switch (en.mode) {
        case VCF_DEL: {                                                     // SVCandidate.py:53-78
            vcf_head(p, e, r.src_tid, s > 1 ? s : 1, en, from);
            if (alleles) {
                const int64_t anchor = s - 1 > 0 ? s - 1 : 0;
                vcf_flush(p, from);
                vcf_ref(p, e, r.src_tid, anchor, t);
                vcf_txt(p, "\t");
                vcf_flush(p, from);
                vcf_ref(p, e, r.src_tid, anchor, s);
            } else {
                vcf_txt(p, "N\t<DEL>");
            }
            vcf_tail_open(p, "PASS");
            vcf_txt(p, "DEL;");
            vcf_end_svlen(p, t, s - t);
            vcf_tail_close(p, r, from);
            break;
        }
        case VCF_INV: {                                                     // SVCandidate.py:99-125
            vcf_head(p, e, r.src_tid, s + 1, en, from);
            if (alleles) {
                vcf_flush(p, from);
                vcf_ref(p, e, r.src_tid, s, t);
                vcf_txt(p, "\t");
                vcf_flush(p, from);
                vcf_ref(p, e, r.src_tid, s, t, 1, true);
            } else {
                vcf_txt(p, "N\t<INV>");
            }
            vcf_tail_open(p, (r.flags & SVB_F_COMPLETE) ? "PASS" : "incomplete_inversion");
            vcf_txt(p, "INV;END=");
            vcf_int(p, t);
            vcf_tail_close(p, r, from);
            break;
        }
        case VCF_INS: {                                                     // SVCandidate.py:151-176 (END is the start)
            vcf_head(p, e, r.dst_tid, ds > 1 ? ds : 1, en, from);
            if (alleles) {
                const int64_t anchor = ds - 1 > 0 ? ds - 1 : 0;
                vcf_flush(p, from);
                vcf_ref(p, e, r.dst_tid, anchor, ds);
                vcf_txt(p, "\t");
                vcf_flush(p, from);
                vcf_ref(p, e, r.dst_tid, anchor, ds);
                vcf_push(p, VP_SEQ4, r.seq_pos, r.seq_len, 0, r.aln_idx, r.hap);
            } else {
                vcf_txt(p, "N\t<INS>");
            }
            vcf_tail_open(p, "PASS");
            vcf_txt(p, "INS;");
            vcf_end_svlen(p, ds, dt - ds);
            vcf_tail_close(p, r, from);
            break;
        }
        case VCF_TAN_AS_INS: {                                              // SVCandidate.py:203-226
            vcf_head(p, e, r.src_tid, s + 1, en, from);
            if (alleles) {
                vcf_flush(p, from);
                vcf_ref(p, e, r.src_tid, s, t);
                vcf_txt(p, "\t");
                vcf_flush(p, from);
                vcf_ref(p, e, r.src_tid, s, t, r.copies + 1 > 0 ? static_cast<uint32_t>(r.copies + 1) : 0u);
            } else {
                vcf_txt(p, "N\t<INS>");
            }
            vcf_tail_open(p, (r.flags & SVB_F_FULLY_COVERED) ? "PASS" : "not_fully_covered");
            vcf_txt(p, "INS;");
            vcf_end_svlen(p, t, (t - s) * r.copies);
            vcf_tail_close(p, r, from);
            break;
        }
        case VCF_TAN_AS_DUP: {                                              // SVCandidate.py:228-261
            vcf_head(p, e, r.src_tid, s + 1, en, from);
            vcf_txt(p, "N\t<DUP:TANDEM>");
            vcf_tail_open(p, (r.flags & SVB_F_FULLY_COVERED) ? "PASS" : "not_fully_covered");
            vcf_txt(p, "DUP:TANDEM;");
            vcf_end_svlen(p, t, t - s);
            vcf_txt(p, "\tGT:CN\t");
            vcf_txt(p, vcf_genotype(r.genotype));
            vcf_txt(p, ":");
            vcf_int(p, static_cast<int64_t>(r.copies) + 1);
            vcf_txt(p, "\n");
            vcf_flush(p, from);
            break;
        }
        case VCF_INT_AS_INS: {                                              // SVCandidate.py:296-322
            vcf_head(p, e, r.dst_tid, ds > 1 ? ds : 1, en, from);
            if (alleles) {
                const int64_t anchor = ds - 1 > 0 ? ds - 1 : 0;
                vcf_flush(p, from);
                vcf_ref(p, e, r.dst_tid, anchor, ds);
                vcf_txt(p, "\t");
                vcf_flush(p, from);
                vcf_ref(p, e, r.dst_tid, anchor, ds);
                vcf_ref(p, e, r.src_tid, s, t);
            } else {
                vcf_txt(p, "N\t<INS>");
            }
            vcf_tail_open(p, "PASS");
            vcf_txt(p, "INS;");
            if (r.flags & SVB_F_CUTPASTE) vcf_txt(p, "CUTPASTE;");
            vcf_end_svlen(p, ds, dt - ds);
            vcf_tail_close(p, r, from);
            break;
        }
        case VCF_INT_AS_DUP: {                                              // SVCandidate.py:324-347
            vcf_head(p, e, r.src_tid, s + 1, en, from);
            vcf_txt(p, "N\t<DUP:INT>");
            vcf_tail_open(p, "PASS");
            vcf_txt(p, "DUP:INT;");
            if (r.flags & SVB_F_CUTPASTE) vcf_txt(p, "CUTPASTE;");
            vcf_end_svlen(p, t, t - s);
            vcf_tail_close(p, r, from);
            break;
        }
        default: {                                                          // BND and its mate, SVCandidate.py:389-443
            const bool mate = en.mode == VCF_BND_MATE;
            const int32_t here_tid = mate ? r.dst_tid : r.src_tid, there_tid = mate ? r.src_tid : r.dst_tid;
            const int64_t here = mate ? ds : s, there = mate ? s : ds;
            const bool sf = (r.flags & SVB_F_SRC_FWD) != 0, df = (r.flags & SVB_F_DST_FWD) != 0;
            // forward record: ff N[p[  fr N]p]  rr ]p]N  rf [p[N ; mate: rr N[p[  fr N]p]  ff ]p]N  rf [p[N
            int shape;                                                      // 0 N[p[  1 N]p]  2 ]p]N  3 [p[N
            if (sf && !df) shape = 1;
            else if (!sf && df) shape = 3;
            else shape = (sf != mate) ? 0 : 2;
            vcf_head(p, e, here_tid, here + 1, en, from);
            vcf_txt(p, "N\t");
            vcf_txt(p, shape == 0 ? "N[" : (shape == 1 ? "N]" : (shape == 2 ? "]" : "[")));
            vcf_flush(p, from);
            vcf_name(p, e, there_tid);
            vcf_txt(p, ":");
            vcf_int(p, there + 1);
            vcf_txt(p, shape == 0 ? "[" : (shape == 1 ? "]" : (shape == 2 ? "]N" : "[N")));
            vcf_tail_open(p, "PASS");
            vcf_txt(p, "BND");
            vcf_tail_close(p, r, from);
            break;
        }
    }
}

SVB_HD uint64_t vcf_plan_length(const VcfPlan& p) {
    uint64_t n = 0;
    for (uint32_t k = 0; k < p.n_pieces; ++k) n += p.piece[k].len;
    return n;
}

// ---- bytes ------------------------------------------------------------------------------------------------------
SVB_HD uint8_t vcf_complement(uint8_t b) {                       // complement.get(base, base), SVCandidate.py:96,106
    return b == 'A' ? 'T' : (b == 'T' ? 'A' : (b == 'C' ? 'G' : (b == 'G' ? 'C' : b)));
}
SVB_HD uint8_t vcf_piece_byte(const VcfPlan& p, const VcfPiece& q, const VcfEnv& e, uint64_t i) {
    switch (q.kind) {
        case VP_TEXT: return static_cast<uint8_t>(p.text[q.src + i]);
        case VP_NAME: return e.names[q.src + i];
        case VP_REF: return e.bases[q.src + (q.len > q.unit ? i % q.unit : i)];
        case VP_REF_REVCOMP: return vcf_complement(e.bases[q.src + (q.len - 1ull - i)]);
        default: {
            const uint64_t nib = e.seq_off[q.slot][q.aux] * 2ull + q.src + i;
            const uint8_t b = e.seq4[q.slot][nib >> 1];
            return static_cast<uint8_t>("=ACMGRSVTWYHKDBN"[(nib & 1ull) ? (b & 15u) : (b >> 4)]);
        }
    }
}
