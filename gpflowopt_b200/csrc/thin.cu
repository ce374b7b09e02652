// Latency path for a handful of candidates (the N = 1 polish calls of a BO iteration; reference call site optim.py:252-276
// through acquisition.py:256-257): contraction kernels shaped for ONE tile of <= 8 candidates instead of thousands.
//
// With so few candidates the contraction is a matrix-vector product: no reuse of the L^-1 / K^-1 stream, so the time is the
// time to pull the stream (16.8 MB + 33.5 MB at M = 2048, L2-resident after the first call) through the SMs' L2 ports.  The
// row-split form of the DMMA kernel gave every CTA 128 rows (2 MB of K^-1 through one SM, ~17 us at the ~120 GB/s one SM can
// pull) and regenerated K(X*,X) for all 8 tile slots in every CTA.  Here a CTA takes 16 rows (two 8-row blocks of the same
// block stream), so one tile spreads over up to all SMs; its 16 warps split the COLUMN blocks of those rows and add their
// partial rows in shared memory in a fixed order; every CTA works on ONE candidate (K(x*,X) generated for it alone -- the
// row-split kernels pad a call to a tile of 8), and the arithmetic is plain DFMA on the DMMA-fragment layout (lane l of a block
// holds row l/4, columns l%4 and 4 + l%4): 2 DFMA per 16 bytes, far below the L2 port rate.  A call of up to four candidates
// is a grid of N x (CTAs per candidate).  (A tile of 4 / 8 candidates per CTA was measured as well: every CTA then generates
// the whole panel for all of them, N = 8 gradients at M = 2048 0.168 ms against 0.114 ms with the row-split kernels.)
// The kernels leave per-CTA partial sums; thin_final_kernel (contract2_aux.cu) adds them in CTA order.
#include <cstdlib>
#include "contract_common.cuh"

constexpr int TH_RB = 2;            // row blocks per unit of work (one CTA pass): 16 rows
constexpr int TH_NW = 16;           // warps per CTA
constexpr int TH_MAX_COLS = 16384;  // training points whose K(x*,X) fits the shared-memory panel (128 KB)
constexpr int TH_U = 4;             // column blocks in flight per warp (x TH_RB 16-byte loads per thread)

template <int KIND, bool GRAD>
__global__ void __launch_bounds__(32 * TH_NW, 2)      // two CTAs per SM: the forward and the gradient launch of a call run side by side
thin_contract_kernel(const GpDev* __restrict__ gpp, EvalIO io) {
    constexpr int RB = TH_RB, ROWS = 8 * RB, THREADS = 32 * TH_NW;
    extern __shared__ __align__(16) double smem[];
    const GpDev& gp = *gpp;
    const int D = gp.D, M = gp.M, R = gp.R;
    const double variance = gp.variance;
    const double* __restrict__ Xs = gp.Xs;
    const double* __restrict__ XsT = gp.XsT;
    const int ldT = gp.ldT;
    const double* __restrict__ xn = gp.xn;
    const PackGeom geom = io.geom_o;
    double* kx = smem;                                    // [8 nrb]  K(x*, X)
    double* red = kx + 8 * geom.nrb;                      // [warp][ROWS] partial rows
    double* xs = red + TH_NW * ROWS;                      // [D] scaled candidate
    double* x2 = xs + D;                                  // [2]
    double* tots = x2 + 2;                                // [1 + RMAX (+1)] running colsum(A^2), A^T V of this CTA's rows
    double* tab32 = tots + 2 + RMAX;                      // [34]
    double* ub = tab32 + 34;                              // GRAD: [1 + RMAX][ROWS]  w dk, alpha_ro dk
    double* gsum = ub + (1 + RMAX) * ROWS;                // GRAD: [1 + R][1 + D]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double* __restrict__ packed = io.packed_o;
    // one candidate per CTA: CTAs [c nsplit, (c + 1) nsplit) work on candidate c of the call
    const int nsplit = io.nsplit, cand = (int)(blockIdx.x / nsplit), split = (int)(blockIdx.x - cand * nsplit);
    const int tile = cand >> 3, ln = cand & 7;                        // its place in the partial-sum layout of the row-split kernels
    const int nunits = (geom.nrb + RB - 1) / RB;

    if (tid < 32) tab32[tid] = exp2(tid * (1.0 / 32.0));
    for (int d = tid; d < D; d += THREADS) {
        const double xt = __dadd_rn(__dmul_rn(io.X[(size_t)cand * D + d], gp.in_scale[d]), gp.in_shift[d]);
        xs[d] = xt / gp.ell[d];
    }
    if (GRAD) for (int e = tid; e < (1 + R) * (1 + D); e += THREADS) gsum[e] = 0.0;
    if (tid < 1 + RMAX) tots[tid] = 0.0;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int d = 0; d < D; ++d) s += xs[d] * xs[d];
        x2[0] = s;
    }
    __syncthreads();

    // K(x*, X) of the column blocks this CTA needs, once (the forms of contract2_kernel's gen_group, same order of the D
    // products).  Thread = training point over the TRANSPOSED inputs: every load is one coalesced 256-byte request per warp
    // (row-major rows of D doubles per lane thrash L1: 40-70 us of a call at M = 4096, D = 16 -- profiles/r02_thin_phase_stamps.log);
    // GJ points per thread at a time keep GJ * 4 independent loads in flight.
    const int u_last = split + (nunits - 1 - split) / nsplit * nsplit;
    const int kp_max = geom.dense ? geom.nrb : (RB * u_last + RB < geom.nrb ? RB * u_last + RB : geom.nrb);
    {
        const int ncol = 8 * kp_max;
        constexpr int GJ = 4;
        const double x2c = x2[0];
#pragma unroll 1
        for (int j0 = tid; j0 < ncol; j0 += GJ * THREADS) {
            double dot[GJ];
            int jc[GJ];
#pragma unroll
            for (int q = 0; q < GJ; ++q) {
                const int col = j0 + q * THREADS;
                jc[q] = col < M ? col : M - 1;
                dot[q] = 0.0;
            }
#pragma unroll 4
            for (int d = 0; d < D; ++d) {
                const double c = xs[d];
#pragma unroll
                for (int q = 0; q < GJ; ++q) dot[q] = fma(__ldg(XsT + (size_t)d * ldT + jc[q]), c, dot[q]);
            }
#pragma unroll
            for (int q = 0; q < GJ; ++q) {
                const int col = j0 + q * THREADS;
                if (col < ncol) {
                    const double r2 = (-2.0 * dot[q] + __ldg(xn + jc[q])) + x2c;
                    const double val = kx_branchless<KIND>(variance, r2, tab32);
                    kx[col] = col < M ? val : 0.0;
                }
            }
        }
    }
    __syncthreads();

    const int frow = tid & (ROWS - 1);                    // finishing threads: tid < ROWS owns row tid of the unit

    for (int u = split; u < nunits; u += nsplit) {
        const int rb0 = RB * u, p = rb0 / geom.rbc, rbl0 = rb0 - p * geom.rbc, nr = pg_chunk_rows(geom, p);
        const int kp_end = geom.dense ? geom.nrb : (rb0 + RB < p * geom.rbc + nr ? rb0 + RB : p * geom.rbc + nr);
        const double* __restrict__ cb = packed + pg_chunk_base(geom, p) * 64 + 2 * lane;
        double acc[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = 0.0;
        // warp w takes column blocks w, w + 16, ...: TH_U of them (TH_U * RB independent 16-byte loads per lane) at a time
        for (int kp0 = w; kp0 < kp_end; kp0 += TH_NW * TH_U) {
            double2 a[TH_U][RB];
#pragma unroll
            for (int q = 0; q < TH_U; ++q) {
                const int kp = kp0 + TH_NW * q;
#pragma unroll
                for (int r = 0; r < RB; ++r) a[q][r] = make_double2(0.0, 0.0);
                if (kp < kp_end) {
                    int lo;
                    const int64_t cbs = pg_col_base(geom, p, nr, kp, &lo);
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        const int rbl = rbl0 + r;
                        if (rbl >= lo && rbl < nr) a[q][r] = ldg_stream(cb + (cbs + rbl - lo) * 64);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < TH_U; ++q) {
                const int kp = kp0 + TH_NW * q;
                if (kp < kp_end) {
                    const double k0 = kx[8 * kp + (lane & 3)], k1 = kx[8 * kp + 4 + (lane & 3)];
#pragma unroll
                    for (int r = 0; r < RB; ++r) acc[r] = fma(a[q][r].x, k0, fma(a[q][r].y, k1, acc[r]));
                }
            }
        }
        // partial rows of this warp: the four lanes of a row add up, then every row is summed over the warps in warp order
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            double v = acc[r];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if ((lane & 3) == 0) red[w * ROWS + r * 8 + (lane >> 2)] = v;
        }
        __syncthreads();
        if (w == 0) {
            double A = 0.0;
            const int rbl = rbl0 + frow / 8;
            const bool rv = lane < ROWS && rbl < nr;
            const int gi = 8 * (p * geom.rbc + rbl) + (frow & 7);       // row of the stream = training point
            if (rv)
                for (int ww = 0; ww < TH_NW; ++ww) A += red[ww * ROWS + frow];
            if (!GRAD) {
                double ssp = A * A, mvp[RMAX];
#pragma unroll
                for (int ro = 0; ro < RMAX; ++ro) mvp[ro] = (rv && ro < R) ? A * __ldg(gp.V + (size_t)gi * R + ro) : 0.0;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {                         // the 16 rows sit in lanes 0 .. 15
                    ssp += __shfl_xor_sync(0xffffffffu, ssp, o);
#pragma unroll
                    for (int ro = 0; ro < RMAX; ++ro) mvp[ro] += __shfl_xor_sync(0xffffffffu, mvp[ro], o);
                }
                if (lane == 0) {
                    tots[0] += ssp;
#pragma unroll
                    for (int ro = 0; ro < RMAX; ++ro) tots[1 + ro] += mvp[ro];
                }
            } else if (lane < ROWS) {
                double dk = 0.0;
                if (rv) {
                    double dot = 0.0;
                    for (int d = 0; d < D; ++d) dot = fma(__ldg(Xs + (size_t)gi * D + d), xs[d], dot);
                    const double r2 = (-2.0 * dot + __ldg(xn + gi)) + x2[0];
                    dk = dk_dr2_branchless<KIND>(variance, r2, tab32);
                }
                ub[frow] = A * dk;
                for (int ro = 0; ro < R; ++ro) ub[(1 + ro) * ROWS + frow] = rv ? __ldg(gp.alpha + (size_t)gi * R + ro) * dk : 0.0;
            }
        }
        __syncthreads();
        if (GRAD) {
            // sum_i u_i [1, Xs_i] over the rows of this unit, every (sum, column) pair owned by one thread
            const int npair = (1 + R) * (1 + D);
            const int rows = 8 * ((nr - rbl0) < RB ? (nr - rbl0) : RB);
            const int g0 = 8 * (p * geom.rbc + rbl0);
            for (int e = tid; e < npair; e += THREADS) {
                const int k = e / (1 + D), c = e - k * (1 + D);
                const double* __restrict__ up = ub + k * ROWS;
                double s = 0.0;
                if (c == 0) for (int row = 0; row < rows; ++row) s += up[row];
                else for (int row = 0; row < rows; ++row) s = fma(up[row], __ldg(Xs + (size_t)(g0 + row) * D + (c - 1)), s);
                gsum[e] += s;
            }
        }
    }
    // partial sums of this CTA in the layouts the row-split kernels use
    if (!GRAD) {
        if (tid < 1 + RMAX) io.partials[(((size_t)tile * nsplit + split) * 8 + ln) * (1 + RMAX) + tid] = tots[tid];   // (written by this thread's warp)
    } else {
        __syncthreads();
        double* part = io.partials + ((size_t)tile * nsplit + split) * (size_t)(8 * (1 + R) * (1 + D));
        for (int e = tid; e < (1 + R) * (1 + D); e += THREADS) {
            const int k = e / (1 + D), c = e - k * (1 + D);
            part[(size_t)(k * 8 + ln) * (1 + D) + c] = gsum[e];
        }
    }
}

template <int KIND, bool GRAD>
static cudaError_t thin_launch_kind(cudaStream_t st, const GpDev* d_gp, EvalIO io, int grid, int D) {
    const size_t smem = ((size_t)8 * io.geom_o.nrb + TH_NW * 8 * TH_RB + (size_t)D + 2 + 2 + RMAX + 34 +
                         (GRAD ? (1 + RMAX) * 8 * TH_RB + (1 + RMAX) * (1 + (size_t)D) : 0)) * sizeof(double);
    auto kern = thin_contract_kernel<KIND, GRAD>;
    static size_t smem_set[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || smem_set[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 16) smem_set[dev] = smem;
    }
    kern<<<grid, 32 * TH_NW, smem, st>>>(d_gp, io);
    return cudaGetLastError();
}

// CTAs of a launch: the 16-row units of the stream spread over the SMs
int gpfo_thin_nsplit(int64_t N, int num_sms, int nrb) {
    (void)N;
    const int nunits = (nrb + TH_RB - 1) / TH_RB;
    return num_sms < nunits ? num_sms : nunits;
}
// Largest call the 16-row kernels take, one candidate per CTA: 4.  Measured (profiles/r02_latency_thin.log, gradients, wall
// clock): M = 2048 N = 2 / 4 / 8: 0.077 / 0.095 / 0.130 ms against 0.108 / 0.115 / 0.121 ms with the row-split DMMA kernels
// (one tile of 8 for all of them); M = 4096, D = 16: 0.130 / 0.193 / 0.318 against 0.226 / 0.222 / 0.222 ms -- every candidate
// pulls the whole stream through the SMs again, so from 8 candidates the shared tile wins.  GPFO_THIN_MAX_N moves it (0: none).
int64_t gpfo_thin_max_n() {
    static const int64_t v = [] { const char* e = getenv("GPFO_THIN_MAX_N"); return e ? atoll(e) : (int64_t)4; }();
    return v;
}
// training points whose K(X*,X) fits the panel of a call (one pass)
int gpfo_thin_panel_cols(int64_t N) { (void)N; return TH_MAX_COLS; }
cudaError_t gpfo_thin_contract_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int D, int grad) {
    const int grid = (int)(io.N * io.nsplit);             // one candidate per CTA
    if (grad) switch (kind) {
        case GPFO_KERN_RBF: return thin_launch_kind<GPFO_KERN_RBF, true>(st, d_gp, io, grid, D);
        case GPFO_KERN_MATERN52: return thin_launch_kind<GPFO_KERN_MATERN52, true>(st, d_gp, io, grid, D);
        default: return thin_launch_kind<GPFO_KERN_MATERN32, true>(st, d_gp, io, grid, D);
    }
    switch (kind) {
        case GPFO_KERN_RBF: return thin_launch_kind<GPFO_KERN_RBF, false>(st, d_gp, io, grid, D);
        case GPFO_KERN_MATERN52: return thin_launch_kind<GPFO_KERN_MATERN52, false>(st, d_gp, io, grid, D);
        default: return thin_launch_kind<GPFO_KERN_MATERN32, false>(st, d_gp, io, grid, D);
    }
}

// Xs (rows x D, row-major) -> XsT (D x rows)
__global__ void transpose_inputs_kernel(const double* __restrict__ Xs, int rows, int D, double* __restrict__ XsT) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)rows * D) return;
    const int d = (int)(e / rows), j = (int)(e - (int64_t)d * rows);
    XsT[e] = Xs[(size_t)j * D + d];
}
void gpfo_transpose_inputs(cudaStream_t st, const double* dXs, int rows, int D, double* dXsT) {
    const int64_t n = (int64_t)rows * D;
    transpose_inputs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dXs, rows, D, dXsT);
}
