// The hot kernel: fused GP prediction + acquisition epilogue + block argmax for one GP over a stream of
// candidates, FP64 throughout, sm_100a.
//
// Replaces, per candidate tile and without ever materialising K(X*, X) in HBM:
//   LinearTransform.build_forward            gpflowopt/transforms.py:102-103   (diagonal affine prologue)
//   kern.K(X, Xnew)                           GPflow 0.5.0 kernels.py           (on-the-fly, expanded form)
//   matrix_triangular_solve(L, Kx)            GPflow 0.5.0 gpr.py               (A = Linv * Kx on DMMA.8x8x4)
//   matmul(A^T, V), Kdiag - reduce_sum(A^2)   GPflow 0.5.0 gpr.py               (register accumulation)
//   build_backward / build_backward_variance  gpflowopt/transforms.py:112-140   (O(1) un-scaling)
//   EI / PoI / PoF / LCB + Sum / Product      gpflowopt/acquisition/*.py        (fused epilogue)
//   np.argmin(evaluations)                    gpflowopt/optim.py:158            (warp-shuffle block argmax)
//
// Work decomposition: a persistent CTA owns candidate tiles of T = 8*NB candidates.  For each tile it walks
// over row chunks of 8*NW*RBW rows of L^-1; warp w owns row blocks {r*NW + w} of the chunk and keeps their
// accumulators (RBW x NB m8n8 tiles) in registers.  L^-1 arrives as a pre-packed stream of 8x8 blocks in
// DMMA lane order (one coalesced 16-byte load per lane per block, software-pipelined one half-column ahead);
// K(X*, X) is generated in groups of KPC column blocks into shared memory in B-fragment order.
#include "contract_common.cuh"

template <int NW, int RBW, int NB>
struct ContractCfg {
    static constexpr int THREADS = 32 * NW;
    static constexpr int RBC = NW * RBW;          // row blocks per chunk
    static constexpr int T = 8 * NB;              // candidates per tile
    static constexpr int HALF = RBW / 2;
    static constexpr int GROUP_SLOTS = KPC * 2 * NB * 32;     // doubles per generated Kx group
    static constexpr int GEN_ITERS = GROUP_SLOTS / THREADS;
    static_assert(NW % NB == 0, "per-thread candidate must be fixed");
    static_assert(GROUP_SLOTS % THREADS == 0, "gen loop must tile");
    static_assert(RBW % 2 == 0, "two halves");
};

template <int NW, int RBW, int NB, int ABL = 0>      // ABL: profiling ablations (1 = no Kx generation, 2 = no L^-1 loads)
__global__ void __launch_bounds__(32 * NW, 1)
contract_kernel(const GpDev* __restrict__ gpp, const ProgramDev* __restrict__ prog, EvalIO io) {
    using C = ContractCfg<NW, RBW, NB>;
    constexpr int T = C::T, HALF = C::HALF;
    extern __shared__ double smem[];
    const int D = gpp->D;
    double* xc = smem;                                   // [D][T] scaled candidate coordinates
    double* x2 = xc + (size_t)D * T;                     // [T]
    double* kx = x2 + T;                                 // [2][GROUP_SLOTS]
    double* red = kx + 2 * C::GROUP_SLOTS;               // [NW][T][1+RMAX]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    const GpDev& gp = *gpp;
    const int M = gp.M, R = gp.R, kind = gp.kind;
    const double variance = gp.variance;
    const double* __restrict__ Xs = gp.Xs;
    const double* __restrict__ xn = gp.xn;
    const double* __restrict__ Vv = gp.V;
    const double* __restrict__ packed = gp.packed;
    const PackGeom geom = gp.geom;

    // generation role of this thread: fixed candidate, fixed (k-step parity / lane) pattern
    const int gen_nb = w % NB;
    const int gen_n = 8 * gen_nb + (lane >> 2);
    const int gen_jl = lane & 3;

    double best_v = 0.0;
    long long best_i = -1;

    const int64_t ntiles = (io.N + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * T;
        __syncthreads();                                  // previous tile's epilogue is done with smem
        // ---- prologue: x' = x*a + b (transforms.py:102-103, diagonal A), then / lengthscales (kernels.py)
        for (int e = tid; e < T * D; e += C::THREADS) {
            const int n = e / D, d = e - n * D;
            double x = 0.0;
            if (n0 + n < io.N) x = io.X[(size_t)(n0 + n) * D + d];
            const double xt = __dadd_rn(__dmul_rn(x, gp.in_scale[d]), gp.in_shift[d]);
            xc[d * T + n] = xt / gp.ell[d];
        }
        __syncthreads();
        if (tid < T) {
            double s = 0.0;
            for (int d = 0; d < D; ++d) s += xc[d * T + tid] * xc[d * T + tid];
            x2[tid] = s;
        }
        __syncthreads();
        const double my_x2 = x2[gen_n];

        double tot_ss = 0.0, tot_mv[RMAX];
#pragma unroll
        for (int ro = 0; ro < RMAX; ++ro) tot_mv[ro] = 0.0;

        for (int p = 0; p < geom.nchunks; ++p) {
            const int nr = pg_chunk_rows(geom, p);
            const int kp_end = p * geom.rbc + nr;        // column blocks this chunk touches: [0, kp_end)
            const double* __restrict__ cbase = packed + pg_chunk_base(geom, p) * 64 + 2 * lane;

            double acc[RBW][NB][2];
#pragma unroll
            for (int r = 0; r < RBW; ++r)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) { acc[r][nb][0] = 0.0; acc[r][nb][1] = 0.0; }

            double2 bufA[HALF], bufB[HALF];
            // loads the blocks {r = H*HALF + rr} of column block kp into buf (inactive ones untouched)
            auto load_half = [&](double2 (&buf)[HALF], int kp, int H) {
                if (kp >= kp_end) return;
                int lo;
                const int64_t cb = pg_col_base(geom, p, nr, kp, &lo);
                const double* __restrict__ src = cbase + (cb - lo) * 64;
#pragma unroll
                for (int rr = 0; rr < HALF; ++rr) {
                    const int rbl = (H * HALF + rr) * NW + w;
                    if (rbl >= lo && rbl < nr) {
                        if (ABL & 2) buf[rr] = make_double2(1e-3 * lane, 2e-3);
                        else buf[rr] = ldg_stream(src + (int64_t)rbl * 64);
                    }
                }
            };

            // ---- generate the first Kx group, start the A pipeline
            auto gen_group = [&](int g, double* __restrict__ dst) {
#pragma unroll
                for (int q = 0; q < C::GEN_ITERS; ++q) {
                    const int ksl = (w + NW * q) / NB;                     // k-step inside the group
                    const int j = (g * KPC) * 8 + 4 * ksl + gen_jl;        // training point
                    double val = 0.0;
                    if (ABL & 1) val = 1e-3 * j;
                    else if (j < M) {
                        const double* __restrict__ xr = Xs + (size_t)j * D;
                        double dot = 0.0;
#pragma unroll 4
                        for (int d = 0; d < D; ++d) dot += __ldg(xr + d) * xc[d * T + gen_n];
                        const double r2 = (-2.0 * dot + __ldg(xn + j)) + my_x2;   // square_dist, expanded form
                        val = kx_from_r2(kind, variance, r2);
                    }
                    dst[(ksl * NB + gen_nb) * 32 + lane] = val;
                }
            };
            const int ngroups = (kp_end + KPC - 1) / KPC;
            gen_group(0, kx);
            load_half(bufA, 0, 0);
            __syncthreads();

            for (int g = 0; g < ngroups; ++g) {
                if (g + 1 < ngroups) gen_group(g + 1, kx + ((g + 1) & 1) * C::GROUP_SLOTS);
                const double* __restrict__ kb = kx + (g & 1) * C::GROUP_SLOTS;
#pragma unroll 1
                for (int kpi = 0; kpi < KPC; ++kpi) {
                    const int kp = g * KPC + kpi;
                    if (kp >= kp_end) break;
                    const int lo = (kp - p * geom.rbc) > 0 ? (kp - p * geom.rbc) : 0;
                    double bf[2][NB];
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) bf[h][nb] = kb[((kpi * 2 + h) * NB + nb) * 32 + lane];
                    load_half(bufB, kp, 1);
#pragma unroll
                    for (int rr = 0; rr < HALF; ++rr) {
                        const int rbl = rr * NW + w;
                        if (rbl >= lo && rbl < nr) {
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb) dmma884(acc[rr][nb][0], acc[rr][nb][1], bufA[rr].x, bf[0][nb]);
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb) dmma884(acc[rr][nb][0], acc[rr][nb][1], bufA[rr].y, bf[1][nb]);
                        }
                    }
                    load_half(bufA, kp + 1, 0);
#pragma unroll
                    for (int rr = 0; rr < HALF; ++rr) {
                        const int rbl = (HALF + rr) * NW + w;
                        if (rbl >= lo && rbl < nr) {
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb)
                                dmma884(acc[HALF + rr][nb][0], acc[HALF + rr][nb][1], bufB[rr].x, bf[0][nb]);
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb)
                                dmma884(acc[HALF + rr][nb][0], acc[HALF + rr][nb][1], bufB[rr].y, bf[1][nb]);
                        }
                    }
                }
                __syncthreads();
            }

            // ---- chunk reduction: colsum(A^2) and A^T V over this chunk's rows (fixed order -> deterministic)
            {
                double ss[NB][2];
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) { ss[nb][0] = 0.0; ss[nb][1] = 0.0; }
#pragma unroll
                for (int r = 0; r < RBW; ++r)
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
                        ss[nb][0] += acc[r][nb][0] * acc[r][nb][0];
                        ss[nb][1] += acc[r][nb][1] * acc[r][nb][1];
                    }
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        double v = ss[nb][c];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        if (lane < 4) red[(w * T + nb * 8 + 2 * lane + c) * (1 + RMAX)] = v;
                    }
                for (int ro = 0; ro < R; ++ro) {
                    double mv[NB][2];
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) { mv[nb][0] = 0.0; mv[nb][1] = 0.0; }
#pragma unroll
                    for (int r = 0; r < RBW; ++r) {
                        const int rbl = r * NW + w;
                        double v = 0.0;
                        if (rbl < nr) v = __ldg(Vv + (size_t)(8 * (p * geom.rbc + rbl) + (lane >> 2)) * R + ro);
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) {
                            mv[nb][0] += acc[r][nb][0] * v;
                            mv[nb][1] += acc[r][nb][1] * v;
                        }
                    }
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            double v = mv[nb][c];
                            v += __shfl_xor_sync(0xffffffffu, v, 4);
                            v += __shfl_xor_sync(0xffffffffu, v, 8);
                            v += __shfl_xor_sync(0xffffffffu, v, 16);
                            if (lane < 4) red[(w * T + nb * 8 + 2 * lane + c) * (1 + RMAX) + 1 + ro] = v;
                        }
                }
            }
            __syncthreads();
            if (tid < T) {
                for (int ww = 0; ww < NW; ++ww) {
                    tot_ss += red[(ww * T + tid) * (1 + RMAX)];
#pragma unroll
                    for (int ro = 0; ro < RMAX; ++ro)
                        if (ro < R) tot_mv[ro] += red[(ww * T + tid) * (1 + RMAX) + 1 + ro];
                }
            }
            __syncthreads();
        }

        // ---- epilogue (warp 0): un-scale, optional moment store, acquisition program, block argmax
        if (w == 0) gpfo_tile_epilogue<T>(gp, prog, io, n0, lane, tot_ss, tot_mv, best_v, best_i);
    }
    if (tid == 0 && io.run_program) {
        io.part_val[blockIdx.x] = best_v;
        io.part_idx[blockIdx.x] = best_i;
    }
}

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

// ------------------------------------------------------------------------------------------------
// variants
// ------------------------------------------------------------------------------------------------
//   1: NW=8, RBW=16, NB=2  -> T=16 candidates/tile, 1024-row chunks   (default for large N)
//   2: NW=8, RBW=16, NB=1  -> T=8                                     (small N: more tiles)
//   3: NW=8, RBW=8,  NB=4  -> T=32, 512-row chunks                    (half the L2 traffic per flop)
template <int NW, int RBW, int NB, int ABL = 0>
static cudaError_t launch_variant(cudaStream_t st, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int grid,
                                  int D) {
    using C = ContractCfg<NW, RBW, NB>;
    const size_t smem = ((size_t)D * C::T + C::T + 2 * C::GROUP_SLOTS + (size_t)NW * C::T * (1 + RMAX)) * sizeof(double);
    auto kern = contract_kernel<NW, RBW, NB, ABL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, C::THREADS, smem, st>>>(d_gp, d_prog, io);
    return cudaGetLastError();
}

// variant table.  1-3 (+ablations 1x/3x): v1 register-prefetch kernel above; 4-6 (+ablations 4x): v2 cp.async kernel
// (contract2.cu).  Ablation results are meaningless numerically.
struct VarInfo { int gen, rbc, tile; };
static VarInfo var_info(int v) {
    if (v == 4 || v == 9 || v == 11 || (v >= 40 && v <= 49)) return {2, 64, 32};
    if (v == 7 || v == 8) return {2, 128, 16};
    if (v == 12 || v == 13 || (v >= 120 && v <= 129)) return {2, 32, 64};
    if (v == 5) return {2, 64, 16};
    if (v == 6) return {2, 64, 8};
    if (v == 3 || (v >= 30 && v <= 39)) return {1, 64, 32};
    if (v == 2) return {1, 128, 8};
    return {1, 128, 16};
}
int gpfo_contract_rbc(int variant) { return var_info(variant).rbc; }
int gpfo_contract_tile(int variant) { return var_info(variant).tile; }

int gpfo_contract_grid(int variant, int64_t N, int num_sms) {
    const int T = gpfo_contract_tile(variant);
    const int64_t ntiles = (N + T - 1) / T;
    return (int)(ntiles < num_sms ? (ntiles > 0 ? ntiles : 1) : num_sms);
}

cudaError_t gpfo_contract2_launch(cudaStream_t st, int variant, int kind, const GpDev* d_gp, const ProgramDev* d_prog,
                                  EvalIO io, int grid, int D);

cudaError_t gpfo_contract_launch(cudaStream_t st, int variant, int kind, const GpDev* d_gp, const ProgramDev* d_prog,
                                 EvalIO io, int grid, int D) {
    if (var_info(variant).gen == 2) return gpfo_contract2_launch(st, variant, kind, d_gp, d_prog, io, grid, D);
    switch (variant) {
        case 2: return launch_variant<8, 16, 1>(st, d_gp, d_prog, io, grid, D);
        case 3: return launch_variant<8, 8, 4>(st, d_gp, d_prog, io, grid, D);
        case 11: return launch_variant<8, 16, 2, 1>(st, d_gp, d_prog, io, grid, D);
        case 12: return launch_variant<8, 16, 2, 2>(st, d_gp, d_prog, io, grid, D);
        case 13: return launch_variant<8, 16, 2, 3>(st, d_gp, d_prog, io, grid, D);
        case 31: return launch_variant<8, 8, 4, 1>(st, d_gp, d_prog, io, grid, D);
        case 32: return launch_variant<8, 8, 4, 2>(st, d_gp, d_prog, io, grid, D);
        case 33: return launch_variant<8, 8, 4, 3>(st, d_gp, d_prog, io, grid, D);
        default: return launch_variant<8, 16, 2>(st, d_gp, d_prog, io, grid, D);
    }
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
