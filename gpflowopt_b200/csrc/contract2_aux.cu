// Contraction kernel v2, the launches around the production forward pass: row-split launches and their finish kernels,
// the stored-A gradient pass (forward with A parked, backward over the reversed transposed stream) and the dense K^-1 pass.
#include <cstdlib>
#include "contract2_kernel.cuh"

// Row-split finish, forward: sums the per-split partials of every candidate in split order, then the usual epilogue
// (moments, acquisition program, argmax).  One warp per 32 candidates, T = candidates per tile of the split launch.
__global__ void __launch_bounds__(128)
split_final_kernel(const GpDev* __restrict__ gpp, const ProgramDev* __restrict__ prog, EvalIO io, int T) {
    __shared__ double sv[4];
    __shared__ long long si[4];
    const GpDev& gp = *gpp;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t n0 = (int64_t)blockIdx.x * 128 + 32 * w, n = n0 + lane;
    double ss = 0.0, mv[RMAX];
#pragma unroll
    for (int ro = 0; ro < RMAX; ++ro) mv[ro] = 0.0;
    if (n < io.N) {
        const int64_t tile = n / T;
        const int ln = (int)(n - tile * T);
        for (int s = 0; s < io.nsplit; ++s) {
            const double* p = io.partials + (((size_t)tile * io.nsplit + s) * T + ln) * (1 + RMAX);
            ss += p[0];
#pragma unroll
            for (int ro = 0; ro < RMAX; ++ro) mv[ro] += p[1 + ro];
        }
    }
    double best_v = 0.0;
    long long best_i = -1;
    gpfo_tile_epilogue<32>(gp, prog, io, n0, lane, ss, mv, best_v, best_i);
    if (lane == 0) { sv[w] = best_v; si[w] = best_i; }
    __syncthreads();
    if (threadIdx.x == 0 && io.run_program) {
        double bv = sv[0];
        long long bi = si[0];
        for (int k = 1; k < 4; ++k)
            if (gpfo_better(sv[k], si[k], bv, bi)) { bv = sv[k]; bi = si[k]; }
        io.part_val[blockIdx.x] = bv;
        io.part_idx[blockIdx.x] = bi;
    }
}

// Row-split finish, gradient pass: d acq / d x_d from the summed (sum_i u_i, sum_i u_i Xs_id) partials.
__global__ void __launch_bounds__(128)
split_final_grad_kernel(const GpDev* __restrict__ gpp, EvalIO io, int T, int D) {
    const GpDev& gp = *gpp;
    const int64_t e = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (e >= io.N * D) return;
    const int64_t n = e / D;
    const int d = (int)(e - n * D);
    const int64_t tile = n / T;
    const int ln = (int)(n - tile * T);
    double s0 = 0.0, sd = 0.0;
    for (int s = 0; s < io.nsplit; ++s) {
        const double* p = io.partials + (((size_t)tile * io.nsplit + s) * T + ln) * (size_t)(1 + D);
        s0 += p[0];
        sd += p[1 + d];
    }
    const double xt = __dadd_rn(__dmul_rn(io.X[e], gp.in_scale[d]), gp.in_shift[d]);
    const double g = 2.0 * gp.in_scale[d] / gp.ell[d] * ((xt / gp.ell[d]) * s0 - sd);
    io.grads[e] = io.grad_accumulate ? io.grads[e] + g : g;
}

// ---- row-split launches (few candidates: the rows of the block stream are spread over the SMs) ----------------------
// forward: T = 8 candidates per tile, 128-row chunks; grid = tiles * nsplit, then the finish kernel
// Warps per CTA of the row-split launches = row blocks per chunk of their stream layout (one row block per warp).  Fewer warps
// per CTA give more, smaller CTAs for the same rows (a tile of <= 8 candidates at M = 2048 on 64 SMs instead of 16) -- measured
// SLOWER (profiles/r02_latency_split_nw.log: N = 1 gradients at M = 2048 0.212 / 0.204 / 0.272 ms for 16 / 8 / 4 warps):
// every CTA generates the K(X*,X) panel left of its rows itself, so the panel generation, not the L^-1 stream, bounds a call.
// 16 stays the default; GPFO_SPLIT_NW keeps the experiment reachable.
static int split_nw() {
    static const int nw = [] {
        const char* e = getenv("GPFO_SPLIT_NW");
        const int v = e ? atoi(e) : 16;
        return (v == 4 || v == 8 || v == 16) ? v : 16;
    }();
    return nw;
}
int gpfo_split_rbc() { return split_nw(); }
int gpfo_split_tile() { return 8; }
cudaError_t gpfo_contract_split_launch(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io,
                                       int D) {
    const int64_t ntiles = (io.N + 7) / 8;
    const int grid = (int)(ntiles * io.nsplit);
    cudaError_t e;
    // ring depth of the L^-1 stream: 8 stages (profiles/r02_latency_split_depth.log: N = 1 gradients at M = 4096 0.419 -> 0.390 ms,
    // no change at M <= 2048); GPFO_SPLIT_S=4 restores round 1's depth
    static const int deep = getenv("GPFO_SPLIT_S") ? atoi(getenv("GPFO_SPLIT_S")) : 8;
    if (split_nw() == 16 && deep == 8) e = launch2<16, 1, 1, 8, 16 + 2048>(st, kind, d_gp, d_prog, io, grid, D);
    else switch (split_nw()) {
        case 16: e = launch2<16, 1, 1, 4, 16 + 2048>(st, kind, d_gp, d_prog, io, grid, D); break;
        case 8: e = launch2<8, 1, 1, 4, 16 + 2048>(st, kind, d_gp, d_prog, io, grid, D); break;
        default: e = launch2<4, 1, 1, 4, 16 + 2048>(st, kind, d_gp, d_prog, io, grid, D); break;
    }
    if (e != cudaSuccess) return e;
    split_final_kernel<<<(unsigned)((io.N + 127) / 128), 128, 0, st>>>(d_gp, d_prog, io, 8);
    return cudaGetLastError();
}
// gradient pass: `wide` = 0: T = 8, one row block per warp (its own dense stream layout); 1: T = 32 over the default K^-1 layout
cudaError_t gpfo_contract_grad_split_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int D, int wide) {
    const int T = wide ? 32 : 8;
    const int64_t ntiles = (io.N + T - 1) / T;
    const int grid = (int)(ntiles * io.nsplit);
    cudaError_t e;
    static const int deep = getenv("GPFO_SPLIT_S") ? atoi(getenv("GPFO_SPLIT_S")) : 8;
    if (wide) e = launch2<16, 4, 4, 4, 4 + 16>(st, kind, d_gp, nullptr, io, grid, D);
    else if (split_nw() == 16 && deep == 8) e = launch2<16, 1, 1, 8, 4 + 16 + 2048>(st, kind, d_gp, nullptr, io, grid, D);
    else switch (split_nw()) {
        case 16: e = launch2<16, 1, 1, 4, 4 + 16 + 2048>(st, kind, d_gp, nullptr, io, grid, D); break;
        case 8: e = launch2<8, 1, 1, 4, 4 + 16 + 2048>(st, kind, d_gp, nullptr, io, grid, D); break;
        default: e = launch2<4, 1, 1, 4, 4 + 16 + 2048>(st, kind, d_gp, nullptr, io, grid, D); break;
    }
    if (e != cudaSuccess) return e;
    split_final_grad_kernel<<<(unsigned)((io.N * D + 127) / 128), 128, 0, st>>>(d_gp, io, T, D);
    return cudaGetLastError();
}

// ---- stored-A gradient pass for large batches: forward = the default kernel (T = 64, 256-row chunks, K(X*,X) scratch) that
// also parks A per tile; backward = the same tiles over the reversed transposed stream, all groups read back, M^2 flop.
int gpfo_contract_bwd_rbc() { return 32; }
int gpfo_contract_bwd_tile() { return 64; }
long long gpfo_contract_astore_stride(int nrb) { return (long long)((nrb + KPC - 1) / KPC) * Contract2Cfg<16, 2, 8, S12>::GROUP_SLOTS; }
size_t gpfo_contract_bwd_smem(int D) {
    using C = Contract2Cfg<16, 2, 8, S12>;
    return ((size_t)16 * C::RING_DOUBLES_PER_WARP + 3 * C::GROUP_SLOTS + (size_t)D * C::T + C::T + C::T * (1 + RMAX) +
            2 * ((C::T + 31) / 32) + 34 + C::T * RMAX + C::T + C::T * (1 + (size_t)D)) * sizeof(double);
}
cudaError_t gpfo_contract_astore_launch(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io,
                                        int grid, int D) {
    return launch2<16, 2, 8, S12, 8 + 128>(st, kind, d_gp, d_prog, io, grid, D);
}
cudaError_t gpfo_contract_bwd_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int grid, int D) {
    return launch2<16, 2, 8, S12, 32>(st, kind, d_gp, nullptr, io, grid, D);
}

// gradient pass launcher (dense K^-1 stream; T = 32 candidates per tile, 512-row chunks)
cudaError_t gpfo_contract_grad_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int grid, int D) {
    return launch2<16, 4, 4, 4, 4>(st, kind, d_gp, nullptr, io, grid, D);
}
int gpfo_contract_grad_rbc() { return 64; }
int gpfo_contract_grad_grid(int64_t N, int num_sms) {
    const int64_t ntiles = (N + 31) / 32;
    return (int)(ntiles < num_sms ? (ntiles > 0 ? ntiles : 1) : num_sms);
}
