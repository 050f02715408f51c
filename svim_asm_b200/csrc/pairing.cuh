// Shared declarations of the pairing stage (K6-K9).
#pragma once
#include <string>

#include "common.cuh"
#include "edit_core.cuh"

constexpr int ED_NCLASS = 32;        // symbol classes of the edit-distance kernel (16 nt16 letters + reference extras)
constexpr int PAIR_MAX = 10;         // partitions larger than this are dropped (SVIM_COMBINE.py:126-128,151-152)
constexpr int PAIR_DIST_STRIDE = PAIR_MAX * (PAIR_MAX - 1) / 2;

struct EditJob {
    HapDesc a, b;
    uint32_t out_index;              // slot in the distance array
    uint32_t pad;
};

int build_class_map(const uint8_t* bases, uint64_t n, uint8_t* map256, std::string* why);
int build_class_map_from_seen(const bool* seen, uint8_t* map256, std::string* why);
int launch_edit_distance(svb_ctx* ctx, const EditJob* d_jobs, uint32_t n_jobs, uint64_t max_text_multi_stripe,
                         const uint8_t* d_ref, const uint8_t* d_seq4_a, const uint8_t* d_seq4_b, const uint8_t* d_class_map,
                         double* d_out);
int launch_scan_u32(svb_ctx* ctx, uint32_t* v, uint32_t n, unsigned long long* d_total);
