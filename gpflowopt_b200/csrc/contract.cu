// Small kernels around the contraction (contract2.cu / contract3.cu): the final argmax reduce, the variant table and the
// FP64 pipe microbenchmarks that provide the roofline denominator.  (The first-generation contraction kernel that lived
// here -- register-prefetched LDG, 8 warps x 255 registers, 15.6 TFLOP/s -- is documented in DESIGN.md 4.1 and profiles/.)
#include "contract_common.cuh"

// Final lexicographic (max value, min index) reduce over per-CTA partials; one block.
__global__ void __launch_bounds__(256) argmax_final_kernel(const double* __restrict__ pv,
                                                           const long long* __restrict__ pi, int n,
                                                           double* __restrict__ out_v, long long* __restrict__ out_i) {
    __shared__ double sv[8];
    __shared__ long long si[8];
    double v = 0.0;
    long long i = -1;
    for (int e = threadIdx.x; e < n; e += 256)
        if (gpfo_better(pv[e], pi[e], v, i)) { v = pv[e]; i = pi[e]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (gpfo_better(ov, oi, v, i)) { v = ov; i = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k)
            if (gpfo_better(sv[k], si[k], v, i)) { v = sv[k]; i = si[k]; }
        *out_v = v; *out_i = i;
    }
}

void gpfo_argmax_final(cudaStream_t st, const double* part_val, const long long* part_idx, int nparts,
                       double* out_val, long long* out_idx) {
    argmax_final_kernel<<<1, 256, 0, st>>>(part_val, part_idx, nparts, out_val, out_idx);
}

// variant table of the contraction kernel (contract2.cu); 10 = 2-CTA cluster kernel (contract3.cu, handled by the API).
//    4: T=32, 512-row chunks, K(X*,X) regenerated per chunk (strictly on chip)      5 / 6: T=16 / T=8 (small batches)
//   11: T=32 + per-CTA K(X*,X) scratch       12: T=64, 256-row chunks + scratch (default for large batches)
//   14: 12 with the scratch re-reads as 1-D bulk copies (cp.async.bulk / UBLKCP + mbarrier): same speed, not default
//   13: the tiling of 12, warp-specialised (contract4.cu: 4 producer warps feed 16 DMMA warps through mbarrier stages)
//   20: v5, 16 candidates x 1024-row chunks, resident K(X*,X) panel in shared memory, no scratch (contract5_kernel.cuh)
//   19: v5, 8 candidates x 1024-row chunks (one wave of 8-candidate tiles: 592 < N <= 1184)
//   21: v5, 32 candidates x 512-row chunks
//   41-43, 121-123: ablations (no generation / no L^-1 loads / neither), 127: per-phase cycle counters -- profiling only
//   (10, 13, 14, 4x and 12x exist only in builds with GPFO_BUILD_EXPERIMENTS=1)
struct VarInfo { int rbc, tile; };
static VarInfo var_info(int v) {
    if (v == 20 || (v >= 22 && v <= 27) || v == 30 || v == 32 || v == 34 || v == 36) return {128, 16};
    if (v == 29) return {64, 16};
    if (v == 19) return {128, 8};
    if (v == 21 || v == 28 || v == 31 || v == 33 || v == 35 || (v >= 37 && v <= 39)) return {64, 32};
    if (v == 12 || v == 13 || v == 14 || (v >= 120 && v <= 129)) return {32, 64};
    if (v == 5) return {64, 16};
    if (v == 6) return {64, 8};
    return {64, 32};
}
int gpfo_contract_rbc(int variant) { return var_info(variant).rbc; }
int gpfo_contract_tile(int variant) { return var_info(variant).tile; }

int gpfo_contract_grid(int variant, int64_t N, int num_sms) {
    const int T = gpfo_contract_tile(variant);
    const int64_t ntiles = (N + T - 1) / T;
    const int ctas = variant == 29 ? 2 * num_sms : num_sms;          // 29: 8-warp CTAs, two per SM
    return (int)(ntiles < ctas ? (ntiles > 0 ? ntiles : 1) : ctas);
}

cudaError_t gpfo_contract2_launch(cudaStream_t st, int variant, int kind, const GpDev* d_gp, const ProgramDev* d_prog,
                                  EvalIO io, int grid, int D);

cudaError_t gpfo_contract5_launch(cudaStream_t st, int variant, int kind, const GpDev* d_gp, const ProgramDev* d_prog,
                                  EvalIO io, int grid, int D);
#ifdef GPFO_BUILD_EXPERIMENTS
cudaError_t gpfo_contract_ws_launch(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io,
                                    int grid, int D);
#endif

cudaError_t gpfo_contract_launch(cudaStream_t st, int variant, int kind, const GpDev* d_gp, const ProgramDev* d_prog,
                                 EvalIO io, int grid, int D) {
    if (variant >= 19 && variant <= 39) return gpfo_contract5_launch(st, variant, kind, d_gp, d_prog, io, grid, D);
#ifdef GPFO_BUILD_EXPERIMENTS
    if (variant == 13) return gpfo_contract_ws_launch(st, kind, d_gp, d_prog, io, grid, D);
#else
    if (variant == 13) return cudaErrorInvalidValue;
#endif
    return gpfo_contract2_launch(st, variant, kind, d_gp, d_prog, io, grid, D);
}

// ------------------------------------------------------------------------------------------------
// FP64 pipe microbenchmarks (register-only loops; 16 warps per SM):
//   mode 0: every warp issues mma.sync m8n8k4 (DMMA.8x8x4)          -> the roofline denominator
//   mode 1: every warp issues DFMA                                   -> FP64 vector rate
//   mode 2: two warps of every scheduler issue DMMA, the other two DFMA, concurrently
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_pipe_kernel(int iters, int mode, double* sink) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    const int w = threadIdx.x >> 5;
    const bool use_dmma = mode == 0 || (mode == 2 && ((w >> 2) & 1) == 0);
    if (use_dmma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                c[i][0] = fma(c[i][0], a, b);
                c[i][1] = fma(c[i][1], b, a);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

// out[0] = DMMA TFLOP/s, out[1] = DFMA TFLOP/s for the given mode (best of 5)
int gpfo_fp64_pipes(cudaStream_t st, int num_sms, int mode, double* out) {
    double* sink = nullptr;
    if (cudaMalloc((void**)&sink, sizeof(double)) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, ctas = num_sms * 2;
    fp64_pipe_kernel<<<ctas, 256, 0, st>>>(200, mode, sink);
    double best_ms = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, st);
        fp64_pipe_kernel<<<ctas, 256, 0, st>>>(iters, mode, sink);
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best_ms = -1.0; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    if (best_ms <= 0.0) return -1;
    const double warps = (double)ctas * 8;
    const double dmma_warps = mode == 0 ? warps : (mode == 2 ? warps / 2 : 0.0);
    const double dfma_warps = warps - dmma_warps;
    out[0] = dmma_warps * 16.0 * iters * 512.0 / (best_ms * 1e-3) / 1e12;
    out[1] = dfma_warps * 32.0 * iters * 64.0 / (best_ms * 1e-3) / 1e12;      // 32 DFMA x 32 lanes x 2 flop
    return 0;
}

double gpfo_dmma_peak(cudaStream_t st, int num_sms) {
    double out[2];
    return gpfo_fp64_pipes(st, num_sms, 0, out) == 0 ? out[0] : -1.0;
}
