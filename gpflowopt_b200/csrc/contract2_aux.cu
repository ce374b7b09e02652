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
static cudaError_t split_forward(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int D) {
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
    return e;
}
cudaError_t gpfo_contract_split_launch(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io,
                                       int D) {
    const cudaError_t e = split_forward(st, kind, d_gp, d_prog, io, D);
    if (e != cudaSuccess) return e;
    split_final_kernel<<<(unsigned)((io.N + 127) / 128), 128, 0, st>>>(d_gp, d_prog, io, 8);
    return cudaGetLastError();
}

// ---- latency path for a handful of candidates (the N = 1 polish calls of a BO iteration): the forward launch of every GP
// and -- for gradients -- the dense K^-1 launch of every GP are independent of each other once the program's partials are
// applied AFTER the contraction (RAWP mode), so they run side by side on two streams; one finish kernel then does what
// split_final_kernel, program_grad_kernel, split_final_grad_kernel and argmax_final_kernel did in four launches.
cudaError_t gpfo_thin_forward_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int D) {
    return split_forward(st, kind, d_gp, nullptr, io, D);
}
cudaError_t gpfo_thin_grad_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int D) {
    const int64_t ntiles = (io.N + 7) / 8;
    return launch2<16, 1, 1, 8, 4 + 16 + 2048 + 4096>(st, kind, d_gp, nullptr, io, (int)(ntiles * io.nsplit), D);
}

// Finish kernel of the latency path, one CTA per candidate.  Everything it needs from global memory -- the partial sums of all
// contraction CTAs, the program, the GPs' scale factors, the candidate -- is requested in ONE round of independent loads (a
// chain of dependent loads costs ~1 us each; the first version of this kernel had ten of them and took 15 us); after that it
// works from shared memory: (1) sums per task in warp order, (2) moments, (3) the acquisition program (warp 7: the value, from
// the instantiation evaluate() uses; the others: its partials in forward mode), (4) d acq / d x by the chain rule through
// every GP, (5) the argmax record.
#define THIN_THREADS 256
__global__ void __launch_bounds__(THIN_THREADS)
thin_final_kernel(const ProgramDev* __restrict__ prog, EvalIO io, ThinFinalArgs a, int ntask) {
    extern __shared__ __align__(16) double psum[];                    // [8 warps][ntask], then [ntask] totals in row 0
    __shared__ ProgramDev sprog;
    __shared__ double svar[GPFO_MAX_GPS], sos[GPFO_MAX_GPS][RMAX], ssh[GPFO_MAX_GPS][RMAX];
    __shared__ double xsd[GPFO_MAX_GPS][GPFO_MAX_DIM], coef[GPFO_MAX_GPS][GPFO_MAX_DIM];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t n = blockIdx.x, tile = n / 8;
    const int ln = (int)(n - tile * 8), D = a.D, ngps = a.ngps;
    // ---- the one round of loads
    for (int e = tid; e < (int)(sizeof(ProgramDev) / sizeof(int)); e += THIN_THREADS)
        reinterpret_cast<int*>(&sprog)[e] = reinterpret_cast<const int*>(prog)[e];
    for (int e = tid; e < ngps * RMAX; e += THIN_THREADS) {
        const int g = e / RMAX, ro = e - g * RMAX;
        if (ro == 0) svar[g] = a.gp[g]->variance;
        if (ro < a.R[g]) { sos[g][ro] = a.gp[g]->out_scale[ro]; ssh[g][ro] = a.gp[g]->out_shift[ro]; }
    }
    if (a.want_grad)
        for (int e = tid; e < ngps * D; e += THIN_THREADS) {
            const int g = e / D, d = e - g * D;
            const GpDev& gp = *a.gp[g];
            const double xt = __dadd_rn(__dmul_rn(io.X[(size_t)n * D + d], gp.in_scale[d]), gp.in_shift[d]);
            xsd[g][d] = xt / gp.ell[d];
            coef[g][d] = 2.0 * gp.in_scale[d] / gp.ell[d];
        }
    // tasks: the (1 + R_g)(1 + D) gradient sums of every GP (only with gradients), then the 1 + RMAX forward sums of every GP;
    // lane = task, warp w adds the contraction CTAs w, w + 8, ... of that task
    const int ntg = ntask - ngps * (1 + RMAX);
    for (int t = lane; t < ntask; t += 32) {
        const double* p;
        size_t stride;
        int ns;
        if (t < ntg) {
            int g = 0, rem = t;
            while (rem >= (1 + a.R[g]) * (1 + D)) { rem -= (1 + a.R[g]) * (1 + D); ++g; }
            const int k = rem / (1 + D), c = rem - k * (1 + D);
            ns = a.ns_g[g];
            stride = (size_t)8 * (1 + a.R[g]) * (1 + D);
            p = a.part_g[g] + (size_t)tile * ns * stride + (size_t)(k * 8 + ln) * (1 + D) + c;
        } else {
            const int g = (t - ntg) / (1 + RMAX), k = (t - ntg) - g * (1 + RMAX);
            ns = a.ns_f[g];
            stride = (size_t)8 * (1 + RMAX);
            p = a.part_f[g] + (size_t)tile * ns * stride + (size_t)ln * (1 + RMAX) + k;
        }
        double s = 0.0;
#pragma unroll 4
        for (int i = w; i < ns; i += THIN_THREADS / 32) s += p[(size_t)i * stride];
        psum[(size_t)w * ntask + t] = s;
    }
    __syncthreads();
    for (int t = tid; t < ntask; t += THIN_THREADS) {
        double s = psum[t];
#pragma unroll
        for (int ww = 1; ww < THIN_THREADS / 32; ++ww) s += psum[(size_t)ww * ntask + t];
        psum[t] = s;
    }
    __syncthreads();
    // ---- moments (as gpfo_tile_epilogue) and the program
    double mu[GPFO_MAX_SLOTS], va[GPFO_MAX_SLOTS];
    for (int g = 0; g < ngps; ++g) {
        const double* mom = psum + ntg + g * (1 + RMAX);
        const double fvar = svar[g] - mom[0];
        for (int ro = 0; ro < a.R[g]; ++ro) {
            mu[a.slot0[g] + ro] = (mom[1 + ro] - ssh[g][ro]) / sos[g][ro];
            va[a.slot0[g] + ro] = fvar / (sos[g][ro] * sos[g][ro]);
        }
    }
    if (w == THIN_THREADS / 32 - 1) {
        const double val = gpfo_run_program<false>(sprog, mu, va, 0.0, nullptr, nullptr, nullptr, nullptr);
        if (lane == 0) {
            if (io.values) io.values[n] = val;
            const bool ok = val == val;                                // NaN never wins (gpfo_better)
            const long long idx = ok ? io.idx0 + n : -1;
            io.part_val[blockIdx.x] = ok ? val : 0.0;
            io.part_idx[blockIdx.x] = idx;
            if (a.best_out && gridDim.x == 1) { a.best_out[0] = ok ? val : 0.0; reinterpret_cast<long long*>(a.best_out)[1] = idx; }
        }
    } else if (a.want_grad && tid < D) {
        double dmu[GPFO_MAX_SLOTS], dva[GPFO_MAX_SLOTS];
        gpfo_run_program<true>(sprog, mu, va, 0.0, nullptr, nullptr, dmu, dva);
        // G_c = cv sum_i w_i dk_i Xa_ic + sum_ro cm_ro sum_i alpha_i,ro dk_i Xa_ic,  Xa = [1, Xs];  d acq / d x_d as in split_final_grad_kernel
        double gacc = 0.0;
        const double* G = psum;
        for (int g = 0; g < ngps; ++g) {
            const int R = a.R[g];
            double cv = 0.0;
            for (int ro = 0; ro < R; ++ro) cv += -2.0 * dva[a.slot0[g] + ro] / (sos[g][ro] * sos[g][ro]);
            double g0 = cv * G[0], gd = cv * G[1 + tid];
            for (int ro = 0; ro < R; ++ro) {
                const double cm = dmu[a.slot0[g] + ro] / sos[g][ro];
                g0 = fma(cm, G[(1 + ro) * (1 + D)], g0);
                gd = fma(cm, G[(1 + ro) * (1 + D) + 1 + tid], gd);
            }
            gacc += coef[g][tid] * (xsd[g][tid] * g0 - gd);
            G += (1 + R) * (1 + D);
        }
        io.grads[(size_t)n * D + tid] = gacc;
    }
}
int gpfo_thin_final_grid(int64_t N) { return (int)N; }
cudaError_t gpfo_thin_final_launch(cudaStream_t st, const ProgramDev* d_prog, EvalIO io, const ThinFinalArgs& a) {
    int ntask = a.ngps * (1 + RMAX);
    if (a.want_grad) for (int g = 0; g < a.ngps; ++g) ntask += (1 + a.R[g]) * (1 + a.D);
    const size_t smem = (size_t)(THIN_THREADS / 32) * ntask * sizeof(double);
    static size_t smem_set[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 40 * 1024 && (dev < 0 || dev >= 16 || smem_set[dev] < smem)) {      // many GPs x many dimensions only
        cudaError_t e = cudaFuncSetAttribute(thin_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 16) smem_set[dev] = smem;
    }
    thin_final_kernel<<<gpfo_thin_final_grid(io.N), THIN_THREADS, smem, st>>>(d_prog, io, a, ntask);
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
