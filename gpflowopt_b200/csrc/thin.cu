// Latency path for a handful of candidates (the N = 1 polish calls of a BO iteration; reference call site optim.py:252-276
// through acquisition.py:256-257): contraction kernels shaped for ONE tile of <= 8 candidates instead of thousands.
//
// With so few candidates the contraction is a matrix-vector product: no reuse of the L^-1 / K^-1 stream, so the time is the
// time to pull the stream (16.8 MB + 33.5 MB at M = 2048, L2-resident after the first call) through the SMs' L2 ports.  The
// row-split form of the DMMA kernel gave every CTA 128 rows (2 MB of K^-1 through one SM, ~17 us at the ~120 GB/s one SM can
// pull) and regenerated K(X*,X) for all 8 tile slots in every CTA.  Here a CTA takes 16 rows (two 8-row blocks of the same
// block stream), so one tile spreads over up to all SMs; its 16 warps split the COLUMN blocks of those rows and add their
// partial rows in shared memory in a fixed order; K(X*,X) is generated for the valid candidates only (NV = 1, 4 or 8), and the
// arithmetic is plain DFMA on the DMMA-fragment layout (lane l of a block holds row l/4, columns l%4 and 4 + l%4), which at
// 2 NV DFMA per 16 bytes is below the L2 port rate even at NV = 8.
// The kernels leave per-CTA partial sums; thin_final_kernel (contract2_aux.cu) adds them in CTA order.
#include <cstdlib>
#include "contract_common.cuh"

constexpr int TH_RB = 2;            // row blocks per unit of work (one CTA pass): 16 rows
constexpr int TH_NW = 16;           // warps per CTA
constexpr int TH_PANEL = 16384;     // doubles of the K(X*,X) panel: 16384 / NV columns per pass
constexpr int TH_U = 4;             // column blocks in flight per warp (x TH_RB 16-byte loads per thread)

// doubles of the K(X*,X) panel: all columns of the stream when they fit, else passes of TH_PANEL / NV columns
__host__ __device__ inline int thin_panel_doubles(int nrb, int nv) { return 8 * nrb * nv < TH_PANEL ? 8 * nrb * nv : TH_PANEL; }

template <int KIND, bool GRAD, int NV>
__global__ void __launch_bounds__(32 * TH_NW, NV == 1 ? 2 : 1)     // N = 1: two CTAs per SM, the forward and the gradient launch of a call run side by side
thin_contract_kernel(const GpDev* __restrict__ gpp, EvalIO io) {
    constexpr int RB = TH_RB, ROWS = 8 * RB, PCB = TH_PANEL / NV / 8, THREADS = 32 * TH_NW;
    extern __shared__ __align__(16) double smem[];
    const GpDev& gp = *gpp;
    const int D = gp.D, M = gp.M, R = gp.R;
    const double variance = gp.variance;
    const double* __restrict__ Xs = gp.Xs;
    const double* __restrict__ XsT = gp.XsT;
    const int ldT = gp.ldT;
    const double* __restrict__ xn = gp.xn;
    const PackGeom geom = io.geom_o;
    double* kxs = smem;                                   // [columns of the pass][NV]
    double* red = kxs + thin_panel_doubles(geom.nrb, NV); // [warp][ROWS][NV] partial rows
    double* xs = red + TH_NW * ROWS * NV;                 // [NV][D] scaled candidates
    double* x2 = xs + NV * D;                             // [NV]
    double* tab32 = x2 + NV;                              // [34]
    double* ub = tab32 + 34;                              // GRAD: [1 + RMAX][NV][ROWS]  W dk, alpha_ro dk
    double* gsum = ub + (1 + RMAX) * NV * ROWS;           // GRAD: [(1 + R) NV][1 + D]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double* __restrict__ packed = io.packed_o;
    const int nsplit = io.nsplit, split = (int)(blockIdx.x % nsplit);
    const int64_t tile = blockIdx.x / nsplit, n0 = tile * 8;
    const int nunits = (geom.nrb + RB - 1) / RB;

    if (tid < 32) tab32[tid] = exp2(tid * (1.0 / 32.0));
    for (int e = tid; e < NV * D; e += THREADS) {
        const int n = e / D, d = e - n * D;
        const double x = n0 + n < io.N ? io.X[(size_t)(n0 + n) * D + d] : 0.0;
        const double xt = __dadd_rn(__dmul_rn(x, gp.in_scale[d]), gp.in_shift[d]);
        xs[e] = xt / gp.ell[d];
    }
    if (GRAD) for (int e = tid; e < (1 + R) * NV * (1 + D); e += THREADS) gsum[e] = 0.0;
    __syncthreads();
    if (tid < NV) {
        double s = 0.0;
        for (int d = 0; d < D; ++d) s += xs[tid * D + d] * xs[tid * D + d];
        x2[tid] = s;
    }
    __syncthreads();

    // K(X*,X) of column blocks [c0, c1) for the NV candidates (the forms of contract2_kernel's gen_group, same order of the
    // D products).  Thread = training point over the TRANSPOSED inputs: every load is one coalesced 256-byte request per warp
    // (row-major rows of D doubles per lane thrash L1: 40-70 us of a call at M = 4096, D = 16 -- profiles/r02_thin_phase_stamps.log);
    // GJ points per thread at a time keep GJ * 4 independent loads in flight.
    auto gen = [&](int c0, int c1) {
        const int ncol = 8 * (c1 - c0);
        constexpr int GJ = 4;
#pragma unroll 1
        for (int n = 0; n < NV; ++n)
#pragma unroll 1
            for (int j0 = tid; j0 < ncol; j0 += GJ * THREADS) {
                double dot[GJ];
                int jc[GJ];
#pragma unroll
                for (int q = 0; q < GJ; ++q) {
                    const int col = 8 * c0 + j0 + q * THREADS;
                    jc[q] = col < M ? col : M - 1;
                    dot[q] = 0.0;
                }
#pragma unroll 4
                for (int d = 0; d < D; ++d) {
                    const double c = xs[n * D + d];
#pragma unroll
                    for (int q = 0; q < GJ; ++q) dot[q] = fma(__ldg(XsT + (size_t)d * ldT + jc[q]), c, dot[q]);
                }
#pragma unroll
                for (int q = 0; q < GJ; ++q) {
                    const int j = j0 + q * THREADS, col = 8 * c0 + j;
                    if (j < ncol) {
                        const double r2 = (-2.0 * dot[q] + __ldg(xn + jc[q])) + x2[n];
                        const double val = kx_branchless<KIND>(variance, r2, tab32);
                        kxs[(size_t)j * NV + n] = col < M ? val : 0.0;
                    }
                }
            }
    };
    // one panel pass covers every unit of this CTA (always, up to M = 2048 NV... 16384 / NV columns): generate it once
    const int u_last = split + (nunits - 1 - split) / nsplit * nsplit;
    const int kp_max = geom.dense ? geom.nrb : (RB * u_last + RB < geom.nrb ? RB * u_last + RB : geom.nrb);
    const bool single = kp_max <= PCB;
    if (single) { gen(0, kp_max); __syncthreads(); }

    constexpr int FW = (ROWS * NV + 31) / 32;             // finishing warps: thread t < ROWS NV owns (candidate t / ROWS, row t % ROWS)
    const bool fl = tid < ROWS * NV;
    const int fn = fl ? tid / ROWS : 0, frow = tid % ROWS;
    double tot_ss = 0.0, tot_mv[RMAX];
#pragma unroll
    for (int ro = 0; ro < RMAX; ++ro) tot_mv[ro] = 0.0;

    for (int u = split; u < nunits; u += nsplit) {
        const int rb0 = RB * u, p = rb0 / geom.rbc, rbl0 = rb0 - p * geom.rbc, nr = pg_chunk_rows(geom, p);
        const int kp_end = geom.dense ? geom.nrb : (rb0 + RB < p * geom.rbc + nr ? rb0 + RB : p * geom.rbc + nr);
        const double* __restrict__ cb = packed + pg_chunk_base(geom, p) * 64 + 2 * lane;
        double acc[RB][NV];
#pragma unroll
        for (int r = 0; r < RB; ++r)
#pragma unroll
            for (int n = 0; n < NV; ++n) acc[r][n] = 0.0;
        for (int c0 = 0; c0 < kp_end; c0 += PCB) {
            const int c1 = c0 + PCB < kp_end ? c0 + PCB : kp_end;
            if (!single) { __syncthreads(); gen(c0, c1); __syncthreads(); }
            for (int kp0 = c0 + w; kp0 < c1; kp0 += TH_NW * TH_U) {
                double2 a[TH_U][RB];
#pragma unroll
                for (int q = 0; q < TH_U; ++q) {
                    const int kp = kp0 + TH_NW * q;
#pragma unroll
                    for (int r = 0; r < RB; ++r) a[q][r] = make_double2(0.0, 0.0);
                    if (kp < c1) {
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
                    if (kp < c1) {
                        const double* __restrict__ kk = kxs + (size_t)(8 * (kp - c0) + (lane & 3)) * NV;
                        double kx[NV], ky[NV];
                        if (NV == 1) { kx[0] = kk[0]; ky[0] = kk[4 * NV]; }
                        else {
#pragma unroll
                            for (int n = 0; n < NV; n += 2) {
                                const double2 t0 = *reinterpret_cast<const double2*>(kk + n);
                                const double2 t1 = *reinterpret_cast<const double2*>(kk + 4 * NV + n);
                                kx[n] = t0.x; kx[n + 1 < NV ? n + 1 : n] = t0.y;
                                ky[n] = t1.x; ky[n + 1 < NV ? n + 1 : n] = t1.y;
                            }
                        }
#pragma unroll
                        for (int r = 0; r < RB; ++r)
#pragma unroll
                            for (int n = 0; n < NV; ++n) acc[r][n] = fma(a[q][r].x, kx[n], fma(a[q][r].y, ky[n], acc[r][n]));
                    }
                }
            }
        }
        // partial rows of this warp: the four lanes of a row add up, then every row is summed over the warps in warp order
#pragma unroll
        for (int r = 0; r < RB; ++r)
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                double v = acc[r][n];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if ((lane & 3) == 0) red[((size_t)w * ROWS + r * 8 + (lane >> 2)) * NV + n] = v;
            }
        __syncthreads();
        if (w < FW) {
            double A = 0.0;
            const int rbl = rbl0 + frow / 8;
            const bool rv = fl && rbl < nr;
            const int gi = 8 * (p * geom.rbc + rbl) + (frow & 7);       // row of the stream = training point
            if (rv)
                for (int ww = 0; ww < TH_NW; ++ww) A += red[((size_t)ww * ROWS + frow) * NV + fn];
            if (!GRAD) {
                double ssp = A * A, mvp[RMAX];
#pragma unroll
                for (int ro = 0; ro < RMAX; ++ro) mvp[ro] = (rv && ro < R) ? A * __ldg(gp.V + (size_t)gi * R + ro) : 0.0;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {                         // the 16 rows of one candidate sit in one half-warp
                    ssp += __shfl_xor_sync(0xffffffffu, ssp, o);
#pragma unroll
                    for (int ro = 0; ro < RMAX; ++ro) mvp[ro] += __shfl_xor_sync(0xffffffffu, mvp[ro], o);
                }
                tot_ss += ssp;
#pragma unroll
                for (int ro = 0; ro < RMAX; ++ro) tot_mv[ro] += mvp[ro];
            } else if (fl) {
                double dk = 0.0;
                if (rv) {
                    double dot = 0.0;
                    for (int d = 0; d < D; ++d) dot = fma(__ldg(Xs + (size_t)gi * D + d), xs[fn * D + d], dot);
                    const double r2 = (-2.0 * dot + __ldg(xn + gi)) + x2[fn];
                    dk = dk_dr2_branchless<KIND>(variance, r2, tab32);
                }
                ub[(size_t)fn * ROWS + frow] = A * dk;
                for (int ro = 0; ro < R; ++ro)
                    ub[((size_t)(1 + ro) * NV + fn) * ROWS + frow] = rv ? __ldg(gp.alpha + (size_t)gi * R + ro) * dk : 0.0;
            }
        }
        __syncthreads();
        if (GRAD) {
            // sum_i u_i [1, Xs_i] over the rows of this unit, every (sum, candidate, column) triple owned by one thread
            const int ntrip = (1 + R) * NV * (1 + D);
            const int rows = 8 * ((nr - rbl0) < RB ? (nr - rbl0) : RB);
            const int g0 = 8 * (p * geom.rbc + rbl0);
            for (int e = tid; e < ntrip; e += THREADS) {
                const int kn = e / (1 + D), c = e - kn * (1 + D);
                const double* __restrict__ up = ub + (size_t)kn * ROWS;
                double s = 0.0;
                if (c == 0) for (int row = 0; row < rows; ++row) s += up[row];
                else for (int row = 0; row < rows; ++row) s = fma(up[row], __ldg(Xs + (size_t)(g0 + row) * D + (c - 1)), s);
                gsum[e] += s;
            }
        }
    }
    if (!GRAD) {
        if (fl && frow == 0 && n0 + fn < io.N) {
            double* part = io.partials + (((size_t)tile * nsplit + split) * 8 + fn) * (1 + RMAX);
            part[0] = tot_ss;
#pragma unroll
            for (int ro = 0; ro < RMAX; ++ro) part[1 + ro] = tot_mv[ro];
        }
    } else {
        __syncthreads();
        double* part = io.partials + ((size_t)tile * nsplit + split) * (size_t)(8 * (1 + R) * (1 + D));
        const int ntrip = (1 + R) * NV * (1 + D);
        for (int e = tid; e < ntrip; e += THREADS) {
            const int kn = e / (1 + D), c = e - kn * (1 + D);
            const int k = kn / NV, n = kn - k * NV;
            part[(size_t)(k * 8 + n) * (1 + D) + c] = gsum[e];
        }
    }
}

template <int KIND, bool GRAD, int NV>
static cudaError_t thin_launch_nv(cudaStream_t st, const GpDev* d_gp, EvalIO io, int grid, int D) {
    const size_t smem = ((size_t)thin_panel_doubles(io.geom_o.nrb, NV) + TH_NW * 8 * TH_RB * NV + (size_t)NV * D + NV + 34 +
                         (GRAD ? (1 + RMAX) * NV * 8 * TH_RB + (1 + RMAX) * NV * (1 + (size_t)D) : 0)) * sizeof(double);
    auto kern = thin_contract_kernel<KIND, GRAD, NV>;
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
// Only NV = 1 is instantiated.  NV = 4 / 8 were measured against the row-split DMMA kernels behind the same finish kernel
// (profiles/r02_latency_thin.log, M = 2048, gradients): N = 4 0.113 vs 0.115 ms, N = 8 0.168 vs 0.114 ms -- every CTA
// generates the whole NV-wide K(X*,X) panel, which at NV = 8 costs more than the 16-row split saves.
template <int KIND, bool GRAD>
static cudaError_t thin_launch_kind(cudaStream_t st, const GpDev* d_gp, EvalIO io, int grid, int D) {
    return thin_launch_nv<KIND, GRAD, 1>(st, d_gp, io, grid, D);
}

// CTAs per tile of 8 candidates: the 16-row units of the stream spread over the SMs
int gpfo_thin_nsplit(int64_t N, int num_sms, int nrb) {
    const int64_t ntiles = (N + 7) / 8;
    const int nunits = (nrb + TH_RB - 1) / TH_RB;
    int ns = (int)(num_sms / ntiles);
    if (ns < 1) ns = 1;
    return ns < nunits ? ns : nunits;
}
// Largest call the 16-row kernels take: one candidate (GPFO_THIN_MAX_N=0: none, the row-split DMMA kernels take N = 1 too)
int64_t gpfo_thin_max_n() {
    static const int64_t v = [] { const char* e = getenv("GPFO_THIN_MAX_N"); return (e && atoll(e) < 1) ? (int64_t)0 : (int64_t)1; }();
    return v;
}
// training points whose K(X*,X) fits the panel of a call (one pass)
int gpfo_thin_panel_cols(int64_t N) { return N == 1 ? TH_PANEL : 0; }
cudaError_t gpfo_thin_contract_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int D, int grad) {
    const int64_t ntiles = (io.N + 7) / 8;
    const int grid = (int)(ntiles * io.nsplit);
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
