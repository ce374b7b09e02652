// Second kernel of the path: per-candidate sum over Pareto cells of Gaussian CDF products (HVPoI), followed by
// the rest of the acquisition program and the block argmax.
//
// Replaces gpflowopt/acquisition/hvpoi.py:104-140: the Phi table normal.cdf(pf_ext) (N x (P+2) x R), the two
// gather_nd's over bounds.ub / bounds.lb (N x C x R temporaries in the reference), PoI = sum_c prod_j dPhi and
// Hv = sum_c [all_j ub > mu] prod_j (ub - max(lb, mu)).  Here one CTA owns TC candidates, keeps their Phi
// tables and pf_ext in shared memory, and streams the int32 cell tables once per TC candidates.
#include "gpfo_internal.cuh"

#define HV_TC 8            // candidates per CTA tile
#define HV_THREADS 256
#define HV_RMAX 8          // objectives

template <int R>
__global__ void __launch_bounds__(HV_THREADS)
hvpoi_kernel(const ProgramDev* __restrict__ prog, CellsDev cells, EvalIO io, int slot_first) {
    extern __shared__ double sm[];
    const int n_ext = cells.n_ext;
    double* pf = sm;                                   // [n_ext][R]
    double* phi = pf + (size_t)n_ext * R;              // [HV_TC][n_ext][R]
    double* cmu = phi + (size_t)HV_TC * n_ext * R;     // [HV_TC][R]
    double* csd = cmu + HV_TC * R;                     // [HV_TC][R]
    double* red = csd + HV_TC * R;                     // [HV_THREADS/32][HV_TC][2]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    for (int e = tid; e < n_ext * R; e += HV_THREADS) pf[e] = cells.pf_ext[e];

    double best_v = 0.0;
    long long best_i = -1;
    const int64_t ntiles = (io.N + HV_TC - 1) / HV_TC;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * HV_TC;
        __syncthreads();
        if (tid < HV_TC * R) {
            const int t = tid / R, j = tid % R;
            const int64_t n = n0 + t;
            double m = 0.0, v = 1.0;
            if (n < io.N) {
                m = io.mom_mean[(size_t)(slot_first + j) * io.N + n];
                v = io.mom_var[(size_t)(slot_first + j) * io.N + n];
            }
            v = fmax(v, GPFO_STABILITY);               // hvpoi.py:109
            cmu[tid] = m;
            csd[tid] = sqrt(v);
        }
        __syncthreads();
        // Phi[t][i][j] = Normal(mu_tj, sd_tj).cdf(pf_ext[i][j])                           (hvpoi.py:112-113)
        for (int e = tid; e < HV_TC * n_ext * R; e += HV_THREADS) {
            const int t = e / (n_ext * R), ij = e % (n_ext * R), j = ij % R;
            phi[e] = gpfo_ndtr((pf[ij] - cmu[t * R + j]) / csd[t * R + j]);
        }
        __syncthreads();
        double poi[HV_TC], hv[HV_TC];
#pragma unroll
        for (int t = 0; t < HV_TC; ++t) { poi[t] = 0.0; hv[t] = 0.0; }
        for (int64_t c = tid; c < cells.C; c += HV_THREADS) {
            int ub[R], lb[R];
            double ubp[R], lbp[R];
#pragma unroll
            for (int j = 0; j < R; ++j) {
                ub[j] = cells.ub[c * R + j] * R + j;
                lb[j] = cells.lb[c * R + j] * R + j;
                ubp[j] = pf[ub[j]];
                lbp[j] = pf[lb[j]];
            }
#pragma unroll
            for (int t = 0; t < HV_TC; ++t) {
                const double* ph = phi + (size_t)t * n_ext * R;
                double pr = 1.0, ext = 1.0;
                bool valid = true;
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    pr *= (ph[ub[j]] - ph[lb[j]]);                                         // hvpoi.py:116-120
                    const double m = cmu[t * R + j];
                    valid = valid && (ubp[j] > m);                                         // hvpoi.py:130-131
                    ext *= (ubp[j] - fmax(lbp[j], m));                                     // hvpoi.py:133-136
                }
                poi[t] += pr;
                hv[t] += valid ? ext : 0.0;
            }
        }
        // deterministic block reduction: shuffle tree inside the warp, fixed warp order across warps
#pragma unroll
        for (int t = 0; t < HV_TC; ++t) {
            double a = poi[t], b = hv[t];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (lane == 0) { red[(w * HV_TC + t) * 2] = a; red[(w * HV_TC + t) * 2 + 1] = b; }
        }
        __syncthreads();
        if (w == 0) {
            double val = 0.0;
            long long idx = -1;
            const int64_t n = n0 + lane;
            if (lane < HV_TC && n < io.N) {
                double a = 0.0, b = 0.0;
                for (int ww = 0; ww < HV_THREADS / 32; ++ww) {
                    a += red[(ww * HV_TC + lane) * 2];
                    b += red[(ww * HV_TC + lane) * 2 + 1];
                }
                const double leaf = b * a;                                                 // Hv * PoI, hvpoi.py:140
                double mu[GPFO_MAX_SLOTS], va[GPFO_MAX_SLOTS];
                for (int s = 0; s < prog->nslots; ++s) {
                    mu[s] = io.mom_mean[(size_t)s * io.N + n];
                    va[s] = io.mom_var[(size_t)s * io.N + n];
                }
                val = gpfo_run_program<false>(*prog, mu, va, leaf, nullptr, nullptr, nullptr, nullptr);
                if (io.values) io.values[n] = val;
                idx = n;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, val, o);
                const long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
                if (gpfo_better(ov, oi, val, idx)) { val = ov; idx = oi; }
            }
            if (gpfo_better(val, idx, best_v, best_i)) { best_v = val; best_i = idx; }
        }
    }
    if (tid == 0) {
        io.part_val[blockIdx.x] = best_v;
        io.part_idx[blockIdx.x] = best_i;
    }
}

template <int R>
static cudaError_t launch_hv(cudaStream_t st, const ProgramDev* d_prog, const CellsDev& cells, EvalIO io, int grid,
                             int slot_first) {
    const size_t smem = ((size_t)cells.n_ext * R * (1 + HV_TC) + 2 * HV_TC * R + (HV_THREADS / 32) * HV_TC * 2) *
                       sizeof(double);
    auto kern = hvpoi_kernel<R>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, HV_THREADS, smem, st>>>(d_prog, cells, io, slot_first);
    return cudaGetLastError();
}

int gpfo_hvpoi_grid(int64_t N, int num_sms) {
    const int64_t ntiles = (N + HV_TC - 1) / HV_TC;
    const int64_t cap = (int64_t)num_sms * 2;
    return (int)(ntiles < cap ? (ntiles > 0 ? ntiles : 1) : cap);
}

size_t gpfo_hvpoi_smem_bytes(int n_ext, int R) {
    return ((size_t)n_ext * R * (1 + HV_TC) + 2 * HV_TC * R + (HV_THREADS / 32) * HV_TC * 2) * sizeof(double);
}

cudaError_t gpfo_hvpoi_launch(cudaStream_t st, const ProgramDev* d_prog, const CellsDev& cells, EvalIO io, int grid,
                              int slot_first) {
    switch (cells.R) {
        case 2: return launch_hv<2>(st, d_prog, cells, io, grid, slot_first);
        case 3: return launch_hv<3>(st, d_prog, cells, io, grid, slot_first);
        case 4: return launch_hv<4>(st, d_prog, cells, io, grid, slot_first);
        case 5: return launch_hv<5>(st, d_prog, cells, io, grid, slot_first);
        case 6: return launch_hv<6>(st, d_prog, cells, io, grid, slot_first);
        default: return cudaErrorInvalidValue;
    }
}
