// Host-side image of a BAM file (what the ingest hands to the python layer) and the two format helpers the host ingest
// (bam_ingest.cpp) and the device ingest (bam_device.cu) share.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/svimasm_b200.h"

struct svb_bam {
    std::vector<std::string> contig_names;
    std::vector<int32_t> contig_len;
    std::string sort_order;
    std::vector<svb_aln_hdr> hdr;
    std::vector<uint32_t> cigar;
    std::vector<svb_segment> seg;
    std::vector<uint32_t> sa_count;
    std::vector<uint8_t> seq4;
    std::vector<uint64_t> seq_off;
    std::vector<char> names;
    std::vector<uint64_t> name_off;
    std::vector<char> sa_text;            // raw SA:Z values, NUL separated (get_tag("SA") of the seam-level API)
    std::vector<int64_t> sa_text_off;     // per record: offset into sa_text or -1
};


// one BGZF member (an independent raw-deflate stream): where its payload sits in the file, where its bytes go
struct BgzfMember {
    uint64_t in_off, in_len, out_off, out_len;
    uint64_t file_off;               // where the member (its gzip header) starts in the file: the `coffset` of BAI virtual offsets
};
// members with ISIZE 0 (the EOF marker) are skipped; false + *why on a malformed file
// (n_threads > 1: the list through the member headers is walked in pieces, see bam_ingest.cpp)
bool bgzf_member_table(const uint8_t* raw, uint64_t size, std::vector<BgzfMember>* blocks, uint64_t* total_out, std::string* why, int n_threads = 1);
// magic, @HD SO, reference names / lengths -> bam; returns the bytes consumed (the first record starts there),
// -1 on a malformed header, -2 when `avail` bytes do not hold the whole header yet
int64_t bam_parse_header(const uint8_t* p, uint64_t avail, svb_bam* bam, std::string* why);
