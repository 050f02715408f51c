/*
 * svimasm_b200_debug.h -- measurement and test hooks of libsvimasm_b200.so.  NOT part of the drop-in surface
 * (include/svimasm_b200.h): bench.py, tools/ and tests/ use these to time kernels on the library's own stream, to count
 * launches, to read back resident data and to pick kernel variants.  Nothing here has a counterpart in the reference.
 */
#ifndef SVIMASM_B200_DEBUG_H
#define SVIMASM_B200_DEBUG_H

#include "svimasm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Kernel timings accumulated since svb_timing_reset(): CUDA events on the library's stream. */
enum { SVB_K_CIGAR_SCAN = 0, SVB_K_SEGMENT_WALK = 1, SVB_K_MERGE = 2, SVB_K_SORT = 3, SVB_K_EDIT_DISTANCE = 4,
       SVB_K_CLUSTER = 5, SVB_K_SCAN_FINALIZE = 6, SVB_K_VCF = 7, SVB_K_COUNT = 8 };
typedef struct {
    double ms[SVB_K_COUNT];
    uint64_t launches[SVB_K_COUNT];
} svb_timing;

int svb_timing_reset(svb_ctx* ctx);
int svb_timing_get(svb_ctx* ctx, svb_timing* out);       /* synchronises the stream first */
int svb_set_scan_variant(svb_ctx* ctx, int variant);      /* cigar_scan loads: 0 = per-warp TMA bulk-copy ring (default), 1 = LDG.128.nc */
/* Step timing on the library's own stream: record marker `slot` (0..15) now; elapsed ms between two markers
 * (synchronises on the later one).  bench.py brackets its timed region with these. */
int svb_launch_count(svb_ctx* ctx, uint64_t* out);       /* kernels this context has launched so far */
int svb_mark(svb_ctx* ctx, int slot);
int svb_elapsed_ms(svb_ctx* ctx, int slot_begin, int slot_end, double* ms);
/* ms of the context's last svb_bam_open_device call: [0] file read, [1] H2D, [2] inflate kernel, [3] record chase,
 * [4] field + copy kernels, [5] host SA parse + record image, [6] total wall, [7] inflated bytes,
 * [8] resident inflate CTAs per SM, [9] mean clock cycles per BGZF member, [10] members, [11] segments of the index-driven
 * record walk (0: the serial walk ran) */
int svb_bam_device_timings(svb_ctx* ctx, double out[12]);
/* the resident reference back on the host: bases (may be NULL), their count, the 256-entry symbol-class map */
int svb_ref_to_host(svb_ctx* ctx, const svb_ref* ref, uint8_t* bases_dst, uint64_t cap, uint64_t* n_bases, uint8_t* class_map256_dst);
/* the context's last svb_pair: [0] partitions, [1] cross-haplotype pairs (edit-distance jobs), [2] pairs that needed the
 * exact kernel after the thresholded wavefront pass, [3] sum of len(h1) x len(h2) over the pairs (full-table cells) */
/* The BGZF member table of a file built by n_threads host threads (csrc/bam_ingest.cpp: the parallel walk of the device ingest
 * against the serial one); out = {members, inflated bytes, checksum over every member's offsets and sizes}. */
int svb_bgzf_member_table_check(const char* path, int n_threads, uint64_t out[3]);
int svb_pair_stats(svb_ctx* ctx, uint64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* SVIMASM_B200_DEBUG_H */
