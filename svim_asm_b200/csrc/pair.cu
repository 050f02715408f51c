// K6-K9 haplotype pairing (placeholder until the pairing kernels land).
#include "common.cuh"
#include "linkage.cuh"

namespace {
__global__ void cluster_labels_kernel(const double* __restrict__ condensed, const uint32_t* __restrict__ n_points,
                                      const uint64_t* __restrict__ offsets, uint32_t n_problems, double threshold,
                                      int32_t* __restrict__ labels_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_problems) return;
    const int n = static_cast<int>(n_points[i]);
    double work[LINK_MAXN * (LINK_MAXN - 1) / 2];
    int labels[LINK_MAXN];
    const double* src = condensed + offsets[i];
    for (int k = 0; k < n * (n - 1) / 2; ++k) work[k] = src[k];
    link_complete_fcluster(n, work, threshold, labels);
    for (int k = 0; k < LINK_MAXN; ++k) labels_out[static_cast<size_t>(i) * LINK_MAXN + k] = k < n ? labels[k] : 0;
}
}  // namespace

int run_cluster_labels(svb_ctx* ctx, const double* condensed, const uint32_t* n_points, uint32_t n_problems,
                       double threshold, int32_t* labels_out) {
    if (!n_problems) return SVB_OK;
    std::vector<uint64_t> off(n_problems + 1, 0);
    for (uint32_t i = 0; i < n_problems; ++i) {
        if (n_points[i] > LINK_MAXN) return svb_fail(ctx, SVB_ERR_ARG, "svb_cluster_labels: more than 32 points");
        off[i + 1] = off[i] + static_cast<uint64_t>(n_points[i]) * (n_points[i] - 1) / 2;
    }
    double* d_c = nullptr;
    uint32_t* d_n = nullptr;
    uint64_t* d_o = nullptr;
    int32_t* d_l = nullptr;
    SVB_CUDA(ctx, cudaMallocAsync(&d_c, sizeof(double) * std::max<uint64_t>(off[n_problems], 1), ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&d_n, sizeof(uint32_t) * n_problems, ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&d_o, sizeof(uint64_t) * (n_problems + 1), ctx->stream));
    SVB_CUDA(ctx, cudaMallocAsync(&d_l, sizeof(int32_t) * LINK_MAXN * n_problems, ctx->stream));
    if (off[n_problems])
        SVB_CUDA(ctx, cudaMemcpyAsync(d_c, condensed, sizeof(double) * off[n_problems], cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(d_n, n_points, sizeof(uint32_t) * n_problems, cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(ctx, cudaMemcpyAsync(d_o, off.data(), sizeof(uint64_t) * (n_problems + 1), cudaMemcpyHostToDevice, ctx->stream));
    {
        KernelTimer timer(ctx, SVB_K_CLUSTER);
        cluster_labels_kernel<<<(n_problems + 63) / 64, 64, 0, ctx->stream>>>(d_c, d_n, d_o, n_problems, threshold, d_l);
    }
    SVB_CUDA(ctx, cudaGetLastError());
    SVB_CUDA(ctx, cudaMemcpyAsync(labels_out, d_l, sizeof(int32_t) * LINK_MAXN * n_problems, cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(d_c, ctx->stream);
    cudaFreeAsync(d_n, ctx->stream);
    cudaFreeAsync(d_o, ctx->stream);
    cudaFreeAsync(d_l, ctx->stream);
    return SVB_OK;
}

int run_pairing(svb_ctx* ctx, const svb_table*, const svb_table*, const svb_records*, const svb_records*, const svb_ref*,
                const svb_params*, svb_table**) {
    return svb_fail(ctx, SVB_ERR_ARG, "svb_pair: not built yet");
}
int run_edit_distance_strings(svb_ctx* ctx, const uint8_t*, const uint64_t*, const uint8_t*, const uint64_t*, uint32_t, int64_t*) {
    return svb_fail(ctx, SVB_ERR_ARG, "svb_edit_distance: not built yet");
}
