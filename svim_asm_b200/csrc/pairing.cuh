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
    uint32_t out_index;              // slot in the distance arrays (= the job's own index)
    uint32_t pad;
};

// K8a thresholded wavefront kernel (wfa.cu).  Per job it leaves an interval [dist, dist_hi]:
//   dist == dist_hi           the exact distance (it is <= t, or one string is contained in the other)
//   t < dist < dist_hi        far: only bounds are known (length difference or t + 1 below, the longer string above)
//   dist < 0                  unknown (strings too long for the kernel's shared memory): the exact kernel must run
struct WfaArgs {
    const EditJob* jobs;
    const unsigned long long* n_jobs_dev;
    unsigned int* counters;          // [0] next job of stage 0 (long pairs), [1] jobs queued for stage 1, [2] next job of stage 1, [3] next job of stage 0 (the rest)
    uint4* big;                      // stage-1 queue: {job, trimmed prefix, trimmed suffix, -}
    uint32_t big_cap;
    const uint8_t* ref;
    const uint8_t* seq4_a;
    const uint8_t* seq4_b;
    const uint8_t* class_map;
    double* dist;
    double* dist_hi;
    uint32_t t;                      // max_edit_distance
    uint32_t cap_chars;              // set by launch_wfa
    int stage;                       // set by launch_wfa
    uint4* profile;                  // SVB_WFA_PROFILE: per job {trimmed la, trimmed lb, waves | stage << 16, cycles}; else NULL (set by launch_wfa)
};
int launch_wfa(svb_ctx* ctx, WfaArgs a);

int build_class_map(const uint8_t* bases, uint64_t n, uint8_t* map256, std::string* why);
int build_class_map_from_seen(const bool* seen, uint8_t* map256, std::string* why);
// K8b exact kernel over the jobs listed in d_list (indices into d_jobs), *d_n_list of them (read on the device: the host
// does not wait for the count).  Writes dist = dist_hi = the distance.  d_need: device word that receives the longest text
// that did not fit the parked-delta stride (0 if all did); the caller re-runs with a larger ctx->ed_stride.
int launch_edit_distance(svb_ctx* ctx, const EditJob* d_jobs, const uint32_t* d_list, const unsigned long long* d_n_list,
                         uint32_t max_jobs, const uint8_t* d_ref, const uint8_t* d_seq4_a, const uint8_t* d_seq4_b,
                         const uint8_t* d_class_map, double* d_dist, double* d_dist_hi, unsigned long long* d_need);
int launch_scan_u32(svb_ctx* ctx, uint32_t* v, uint32_t n, unsigned long long* d_total);
