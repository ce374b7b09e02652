// Contraction kernel, warp-specialised form (variant 13): the work decomposition and arithmetic of the default kernel
// (contract2.cu, variant 12: 64-candidate tiles, 256-row chunks, K(X*,X) generated once per tile into a per-CTA scratch),
// but K(X*,X) groups are produced by four dedicated warps and consumed by the sixteen DMMA warps through a ring of
// shared-memory stages guarded by mbarriers, instead of every warp alternating between generation and DMMA work under
// a CTA-wide barrier per group:
//   * consumers (warps 0-15, one per row-block slot as before) never execute generation code and never meet a CTA
//     barrier inside a chunk: per group one mbarrier wait (stage full) and one arrive (stage free);
//   * producers (warps 16-19, one per scheduler) run up to KXB-1 groups ahead: a group left of the chunk's diagonal is
//     copied back from the scratch with cp.async (completion signalled by cp.async.mbarrier.arrive), a diagonal group
//     is generated (two adjacent training points x one candidate per step, 16-byte stores to the stage and the scratch);
//     every producer thread writes and later re-reads exactly the same 16-byte scratch chunks, so no fence is needed;
//   * setmaxnreg moves registers from the producer warpgroup (64) to the consumer warpgroups (104).
#include <type_traits>
#include "contract_common.cuh"

namespace {

constexpr int WS_NW = 16, WS_RBW = 2, WS_NB = 8, WS_S = 4;      // consumer tiling = Contract2Cfg<16, 2, 8, 4>
constexpr int WS_PT = 128;                                       // producer threads (one warpgroup)
constexpr int WS_THREADS = 32 * WS_NW + WS_PT;

__device__ __forceinline__ void cp_async_mbar_arrive(unsigned bar_s) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_s) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int KIND, int KXB>
__global__ void __launch_bounds__(WS_THREADS, 1)
contract_ws_kernel(const GpDev* __restrict__ gpp, const ProgramDev* __restrict__ prog, EvalIO io) {
    using C = Contract2Cfg<WS_NW, WS_RBW, WS_NB, WS_S>;
    constexpr int NW = WS_NW, RBW = WS_RBW, NB = WS_NB, S = WS_S, T = C::T;
    extern __shared__ __align__(16) double smem[];
    const int D = gpp->D;
    double* ring = smem;                                             // [NW][S][RBW][64]   consumers: L^-1 blocks in flight
    double* kx = ring + (size_t)NW * C::RING_DOUBLES_PER_WARP;       // [KXB][GROUP_SLOTS] K(X*,X) stages
    double* xc = kx + KXB * C::GROUP_SLOTS;                          // [D][T]             producers only
    double* x2 = xc + (size_t)D * T;                                 // [T]
    double* tot = x2 + T;                                            // [T][1+RMAX]        consumers only
    constexpr int EW = (T + 31) / 32;
    double* sbest = tot + T * (1 + RMAX);                            // [EW][2]
    double* tab32 = sbest + 2 * EW;                                  // [32] 2^(i/32)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tab32 + 32);    // full[KXB], empty[KXB]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    const GpDev& gp = *gpp;
    const int M = gp.M, R = gp.R;
    const PackGeom geom = gp.geom;
    const unsigned bar_full = (unsigned)__cvta_generic_to_shared(bars);
    const unsigned bar_empty = bar_full + 8 * KXB;

    if (tid == 0) {
        for (int s = 0; s < KXB; ++s) { mbar_init(bar_full + 8 * s, WS_PT); mbar_init(bar_empty + 8 * s, NW); }
    }
    if (tid < EW) { sbest[2 * tid] = 0.0; reinterpret_cast<long long*>(sbest)[2 * tid + 1] = -1; }
    if (tid < 32) tab32[tid] = exp2(tid * (1.0 / 32.0));
    __syncthreads();

    const int64_t ntiles = (io.N + T - 1) / T;
    double* const scratch = io.kx_scratch + (size_t)blockIdx.x * io.kx_scratch_stride;

    if (w >= NW) {
        // =============================== producers ===============================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        const int pt = tid - 32 * NW;
        const double variance = gp.variance;
        const double* __restrict__ Xs = gp.Xs;
        const double* __restrict__ xn = gp.xn;
        const int gen_n = 8 * (pt >> 4) + ((pt & 15) >> 1);          // this thread's candidate (fixed)
        const int gen_j = 2 * (pt & 1);                              // ... and its pair of rows within every 4-row slice
        const unsigned kx_s = (unsigned)__cvta_generic_to_shared(kx);
        unsigned it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t n0 = tile * T;
            named_bar_sync(2, WS_PT);                                // every producer is done with the previous tile's xc
            for (int e = pt; e < T * D; e += WS_PT) {
                const int n = e / D, d = e - n * D;
                double x = 0.0;
                if (n0 + n < io.N) {
                    if (io.gen_A) {     // RandomDesign on the device (design.py:55-70, 93-94), as in contract2.cu
                        const double u = gpfo_uniform(io.gen_seed, (unsigned long long)(io.gen_first + n0 + n) * D + d);
                        x = __dadd_rn(__dmul_rn(u, io.gen_A[d]), io.gen_b[d]);
                    } else {
                        x = io.X[(size_t)(n0 + n) * D + d];
                    }
                }
                const double xt = __dadd_rn(__dmul_rn(x, gp.in_scale[d]), gp.in_shift[d]);
                xc[d * T + n] = xt / gp.ell[d];
            }
            named_bar_sync(2, WS_PT);
            if (pt < T) {
                double s = 0.0;
                for (int d = 0; d < D; ++d) s += xc[d * T + pt] * xc[d * T + pt];
                x2[pt] = s;
            }
            named_bar_sync(2, WS_PT);
            const double myx2 = x2[gen_n];
            for (int p = 0; p < geom.nchunks; ++p) {
                const int nr = pg_chunk_rows(geom, p);
                const int row0 = p * geom.rbc;
                const int kp_end = row0 + nr;
                const int ngroups = (kp_end + KPC - 1) / KPC;
                const int nold = row0 / KPC;
                for (int g = 0; g < ngroups; ++g, ++it) {
                    const unsigned stage = it % KXB;
                    unsigned parity = ((it / KXB) & 1) ^ 1;          // a fresh barrier passes the first wait on parity 1
                    mbar_wait(bar_empty + 8 * stage, parity);
                    double* const gsrc = scratch + (size_t)g * C::GROUP_SLOTS;
                    if (g < nold) {
                        const unsigned dst = kx_s + stage * (C::GROUP_SLOTS * 8);
#pragma unroll
                        for (int q = 0; q < C::GROUP_SLOTS / (2 * WS_PT); ++q) {
                            const int ch = pt + q * WS_PT;
                            cp_async16(dst + ch * 16, gsrc + 2 * ch);
                        }
                        cp_async_mbar_arrive(bar_full + 8 * stage);
                    } else {
                        double* const dsts = kx + stage * C::GROUP_SLOTS;
                        // slots 2 (pt + 128 q), +1  <->  rows j = 64 g + 4 q + gen_j, +1 of candidate gen_n; four q per pass
#pragma unroll 1
                        for (int q0 = 0; q0 < C::GROUP_SLOTS / (2 * WS_PT); q0 += 4) {
                            double dot[8];
                            int jj[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int j = (g * KPC) * 8 + 4 * (q0 + (u >> 1)) + gen_j + (u & 1);
                                jj[u] = j < M ? j : M - 1;
                                dot[u] = 0.0;
                            }
                            for (int d = 0; d < D; ++d) {
                                const double c = xc[d * T + gen_n];
#pragma unroll
                                for (int u = 0; u < 8; ++u) dot[u] = fma(__ldg(Xs + (size_t)jj[u] * D + d), c, dot[u]);
                            }
#pragma unroll
                            for (int u = 0; u < 8; u += 2) {
                                double v[2];
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const int j = (g * KPC) * 8 + 4 * (q0 + (u >> 1)) + gen_j + h;
                                    const double r2 = (-2.0 * dot[u + h] + __ldg(xn + jj[u + h])) + myx2;
                                    const double val = kx_branchless<KIND>(variance, r2, tab32);
                                    v[h] = j < M ? val : 0.0;
                                }
                                const int s0 = 2 * (pt + (q0 + (u >> 1)) * WS_PT);
                                *reinterpret_cast<double2*>(dsts + s0) = make_double2(v[0], v[1]);
                                __stcg(reinterpret_cast<double2*>(gsrc + s0), make_double2(v[0], v[1]));
                            }
                        }
                        mbar_arrive(bar_full + 8 * stage);
                    }
                }
            }
        }
        cp_async_wait<0>();
        return;
    }

    // =================================== consumers =====================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const double* __restrict__ Vv = gp.V;
    const double* __restrict__ packed = gp.packed;
    double* myring = ring + (size_t)w * C::RING_DOUBLES_PER_WARP + 2 * lane;
    const unsigned myring_s = (unsigned)__cvta_generic_to_shared(myring);
    unsigned it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * T;
        named_bar_sync(1, 32 * NW);                          // the previous tile's epilogue has read tot
        if (tid < T * (1 + RMAX)) tot[tid] = 0.0;
        for (int p = 0; p < geom.nchunks; ++p) {
            const int nr = pg_chunk_rows(geom, p);
            const int row0 = p * geom.rbc;
            const int kp_end = row0 + nr;
            const int kd0 = row0;
            const double* __restrict__ cbase = packed + pg_chunk_base(geom, p) * 64 + 2 * lane;

            double acc[RBW][NB][2];
#pragma unroll
            for (int r = 0; r < RBW; ++r)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) { acc[r][nb][0] = 0.0; acc[r][nb][1] = 0.0; }

            // L^-1 stream state: running source pointer of column block kp_issue and ONE ring cursor (the write stage is
            // always S-1 stages ahead of the read stage), to keep the loop-carried state small at 104 registers
            const double* __restrict__ src_issue = cbase + (int64_t)w * 64;
            int kp_issue = 0, rd_stage = 0;
            auto issue = [&](int wr_stage) {
                if (kp_issue < kp_end) {
                    const int kd = kp_issue - kd0;
                    const int lo = kd > 0 ? kd : 0;
                    const unsigned dst = myring_s + wr_stage * 8;
#pragma unroll
                    for (int r = 0; r < RBW; ++r) {
                        const int rbl = r * NW + w;
                        if (rbl >= lo && rbl < nr) cp_async16(dst + r * 512, src_issue + r * (NW * 64));
                    }
                    src_issue += (kd < 0 ? nr : nr - kd - 1) * 64;
                }
                cp_async_commit();
                ++kp_issue;
            };
            auto mma_kp = [&](const double* __restrict__ kbk, int lo, auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                issue((rd_stage + (S - 1) * RBW * 64) & (S * RBW * 64 - 1));
                cp_async_wait<S - 1>();
                const double* stage = myring + rd_stage;
                double2 a[RBW];
#pragma unroll
                for (int r = 0; r < RBW; ++r) {
                    const int rbl = r * NW + w;
                    a[r] = (FULL || (rbl >= lo && rbl < nr)) ? *reinterpret_cast<const double2*>(stage + r * 64) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {                  // one half of the B fragments at a time: 16 live registers
                    double bf[NB];
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) bf[nb] = kbk[(h * NB + nb) * 32];
#pragma unroll
                    for (int r = 0; r < RBW; ++r) {
                        const int rbl = r * NW + w;
                        if (FULL || (rbl >= lo && rbl < nr)) {
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb) dmma884(acc[r][nb][0], acc[r][nb][1], h ? a[r].y : a[r].x, bf[nb]);
                        }
                    }
                }
                rd_stage = (rd_stage + RBW * 64) & (S * RBW * 64 - 1);
            };

            const int ngroups = (kp_end + KPC - 1) / KPC;
#pragma unroll
            for (int s = 0; s < S - 1; ++s) issue(s * RBW * 64);
            for (int g = 0; g < ngroups; ++g, ++it) {
                const unsigned stg = it % KXB;
                unsigned parity = (it / KXB) & 1;
                mbar_wait(bar_full + 8 * stg, parity);
                const double* __restrict__ kb = kx + stg * C::GROUP_SLOTS + lane;
                const int k0 = g * KPC;
                const int k1 = (k0 + KPC < kp_end) ? k0 + KPC : kp_end;
                if (nr == C::RBC && k1 <= kd0) {
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp) mma_kp(kb + (kp - k0) * (2 * NB * 32), 0, std::true_type());
                } else {
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp)
                        mma_kp(kb + (kp - k0) * (2 * NB * 32), kp > kd0 ? kp - kd0 : 0, std::false_type());
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * stg);
            }
            cp_async_wait<0>();

            // ---- chunk reduction (fixed order -> deterministic); slices alias the drained per-warp rings
            double* const red = ring;
            constexpr int RS = C::RING_DOUBLES_PER_WARP;
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
                        if (lane < 4) red[w * RS + (nb * 8 + 2 * lane + c) * (1 + RMAX)] = v;
                    }
                for (int ro = 0; ro < R; ++ro) {
                    double mv[NB][2];
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) { mv[nb][0] = 0.0; mv[nb][1] = 0.0; }
#pragma unroll
                    for (int r = 0; r < RBW; ++r) {
                        const int rbl = r * NW + w;
                        double v = 0.0;
                        if (rbl < nr) v = __ldg(Vv + (size_t)(8 * (row0 + rbl) + (lane >> 2)) * R + ro);
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
                            if (lane < 4) red[w * RS + (nb * 8 + 2 * lane + c) * (1 + RMAX) + 1 + ro] = v;
                        }
                }
            }
            named_bar_sync(1, 32 * NW);
            if (tid < T * (1 + RMAX)) {
                const int n = tid / (1 + RMAX), k = tid - n * (1 + RMAX);
                if (k <= R) {
                    double a = tot[tid];
                    for (int ww = 0; ww < NW; ++ww) a += red[ww * RS + n * (1 + RMAX) + k];
                    tot[tid] = a;
                }
            }
            named_bar_sync(1, 32 * NW);
        }
        if (w < EW) {
            constexpr int TE = T < 32 ? T : 32;
            double tot_mv[RMAX];
            const int ln = 32 * w + (lane < TE ? lane : 0);
#pragma unroll
            for (int ro = 0; ro < RMAX; ++ro) tot_mv[ro] = tot[ln * (1 + RMAX) + 1 + ro];
            double best_v = sbest[2 * w];
            long long best_i = reinterpret_cast<long long*>(sbest)[2 * w + 1];
            gpfo_tile_epilogue<TE>(gp, prog, io, n0 + 32 * w, lane, tot[ln * (1 + RMAX)], tot_mv, best_v, best_i);
            if (lane == 0) { sbest[2 * w] = best_v; reinterpret_cast<long long*>(sbest)[2 * w + 1] = best_i; }
        }
    }
    named_bar_sync(1, 32 * NW);
    if (tid == 0 && io.run_program) {
        double bv = sbest[0];
        long long bi = reinterpret_cast<long long*>(sbest)[1];
        for (int k = 1; k < EW; ++k)
            if (gpfo_better(sbest[2 * k], reinterpret_cast<long long*>(sbest)[2 * k + 1], bv, bi)) {
                bv = sbest[2 * k]; bi = reinterpret_cast<long long*>(sbest)[2 * k + 1];
            }
        io.part_val[blockIdx.x] = bv;
        io.part_idx[blockIdx.x] = bi;
    }
}

template <int KIND, int KXB>
cudaError_t launch_ws(cudaStream_t st, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int grid, int D) {
    using C = Contract2Cfg<WS_NW, WS_RBW, WS_NB, WS_S>;
    const size_t smem = ((size_t)WS_NW * C::RING_DOUBLES_PER_WARP + KXB * C::GROUP_SLOTS + (size_t)D * C::T + C::T +
                         C::T * (1 + RMAX) + 2 * ((C::T + 31) / 32) + 32 + 2 * KXB) * sizeof(double);
    auto kern = contract_ws_kernel<KIND, KXB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, WS_THREADS, smem, st>>>(d_gp, d_prog, io);
    return cudaGetLastError();
}

template <int KXB>
cudaError_t launch_ws_kind(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int grid, int D) {
    switch (kind) {
        case GPFO_KERN_RBF: return launch_ws<GPFO_KERN_RBF, KXB>(st, d_gp, d_prog, io, grid, D);
        case GPFO_KERN_MATERN52: return launch_ws<GPFO_KERN_MATERN52, KXB>(st, d_gp, d_prog, io, grid, D);
        default: return launch_ws<GPFO_KERN_MATERN32, KXB>(st, d_gp, d_prog, io, grid, D);
    }
}

}  // namespace

// Four K(X*,X) stages when they fit next to the candidate tile (D <= 32), three otherwise.
cudaError_t gpfo_contract_ws_launch(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io,
                                    int grid, int D) {
    if (D <= 32) return launch_ws_kind<4>(st, kind, d_gp, d_prog, io, grid, D);
    return launch_ws_kind<3>(st, kind, d_gp, d_prog, io, grid, D);
}
