// Backward launch of the stored-A gradient pass in the tall-chunk tiling of contract5_kernel.cuh (16 candidates x 1024-row chunks):
//   W = L^-T A      over the reversed transposed block stream (row i' of the stream = training point rev - i'), A read back
//                   from the scratch the forward launch (contract5_kernel, ABL bit 512) parked it in, 8 KB per column group,
//   u_in = (sum_ro c_mu[n][ro] alpha[i][ro] + c_var[n] w_in) dk_in/dr^2,      G[n][c] = sum_i u_in [1, Xs_i][c],
//   d acq / d x_d = 2 a_d / l_d (xs_d G[n][0] - G[n][1 + d]).
// What differs from the forward kernel:
//   * nothing is generated: a column group of A arrives by ONE bulk copy (cp.async.bulk, the TMA engine; SASS UBLKCP) issued by
//     an elected lane and signalled on the group's mbarrier (expect_tx), into the resident panel (chunk-column 0) or the
//     streaming ring (empty[] barriers as in the forward kernel);
//   * the chunk epilogue stays in registers: the dot products Xs_i . xc_n that dk/dr^2 needs are DMMA tiles in the accumulator
//     layout; u (accumulator layout) becomes a B operand by four warp shuffles per 8 x 8 tile; G^T += Xa^T u runs on the tensor pipe
//     (one DMMA per tile, k-half and column block of Xa = [1, Xs]); per-warp partial tiles are summed in a fixed order.
// (contract2_kernel's backward mode, which this replaces for large batches, walks 64-candidate tiles in 256-row chunks with a CTA
//  barrier per column group and a shared-memory u buffer: 68 % DMMA-active.)
#pragma once
#include "contract5_kernel.cuh"

__device__ __forceinline__ void mbar_expect_tx(unsigned bar_s, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(unsigned dst_s, const void* src, unsigned bytes, unsigned bar_s) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
}

template <int RBW, int NB, int QD>
struct Contract5BwdCfg {
    using F = Contract5Cfg<RBW, NB, QD>;
    static size_t smem_bytes(int D, int nbuf) {
        const int DP = (D + 3) / 4 * 4;
        return ((size_t)F::NW * F::RING_DOUBLES_PER_WARP + (size_t)(F::RG + nbuf) * F::GROUP_SLOTS + 2 * (size_t)DP * F::T +
                2 * F::T * (RMAX + 2) + (size_t)F::T * (1 + D) + 34) * sizeof(double) + (size_t)(F::RG + 2 * F::NBUF_MAX) * 8;
    }
        // the chunk epilogue keeps four column blocks of Xa = [1, Xs] per pass
    static bool fits(int D) { return (1 + D + 7) / 8 <= 4; }
};

template <int RBW, int NB, int QD, int KIND, int ABL = 0>
__global__ void __launch_bounds__(512, 1)
contract5_bwd_kernel(const GpDev* __restrict__ gpp, EvalIO io, int nbuf) {
    using C = Contract5Cfg<RBW, NB, QD>;
    constexpr int T = C::T, NW = C::NW, GK = C::GK, RQ = C::RQ, QPC = C::QPC, RG = C::RG, KA = C::KA;
    constexpr int GS = C::GROUP_SLOTS;
    extern __shared__ __align__(16) double smem[];
    const int D = gpp->D;
    const int DP = (D + 3) / 4 * 4;
    double* ring = smem;                                             // [NW][QD][RQ][64]  private rings of the block stream
    double* kres = ring + (size_t)NW * C::RING_DOUBLES_PER_WARP;     // [RG][GS]          resident panel of A (chunk-column 0)
    double* kstr = kres + (size_t)RG * GS;                           // [nbuf][GS]        streaming groups right of the panel
    double* xcb = kstr + (size_t)nbuf * GS;                          // [2][DP][T]        candidates of this / the next tile
    double* gcb = xcb + 2 * (size_t)DP * T;                          // [2][T][RMAX+2]    c_mu (RMAX), c_var, |xc|^2 per candidate
    double* gsum = gcb + 2 * T * (RMAX + 2);                         // [T][1+D]          G of the current tile
    double* tab32 = gsum + (size_t)T * (1 + D);                      // [32] 2^(i/32) (+2 pad)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tab32 + 34);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int wr = w;                                 // row slot of the warp (contract5_mma.inc); no triangle balancing here
    const GpDev& gp = *gpp;
    const int M = gp.M, R = gp.R;
    const double variance = gp.variance;
    const double* __restrict__ Xs = gp.Xs;
    const double* __restrict__ xn = gp.xn;
    const PackGeom geom = io.geom_o;                  // reversed transposed stream, chunks of RBC row blocks
    const double* __restrict__ packed = io.packed_o;
    double* myring = ring + (size_t)w * C::RING_DOUBLES_PER_WARP + 2 * lane;
    const unsigned myring_s = (unsigned)__cvta_generic_to_shared(myring);
    const unsigned bar_res = (unsigned)__cvta_generic_to_shared(bars);
    const unsigned bar_full = bar_res + 8 * RG, bar_empty = bar_full + 8 * C::NBUF_MAX;
    const unsigned kres_s = (unsigned)__cvta_generic_to_shared(kres), kstr_s = (unsigned)__cvta_generic_to_shared(kstr);
    const int la = nbuf >= 3 ? 2 : 1;                 // groups are requested `la` steps ahead of use
    const int rev = 8 * geom.nrb - 1;                 // stream row i' is training point rev - i'
    const int ND = (1 + D + 7) / 8;                   // column blocks of Xa = [1, Xs]

    if (tid < 32) tab32[tid] = exp2(tid * (1.0 / 32.0));
    if (tid == 0) {
        for (int b = 0; b < RG; ++b) mbar_init(bar_res + 8 * b, 1);                 // one arrival: the lane that issues the copy
        for (int b = 0; b < C::NBUF_MAX; ++b) { mbar_init(bar_full + 8 * b, 1); mbar_init(bar_empty + 8 * b, NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // candidates and chain-rule coefficients of tile `tile` -> xcb[buf], gcb[buf]
    auto stage_tile = [&](int64_t tile, int buf, int first_thread, int nthreads) {
        const int64_t n0 = tile * T;
        double* xc = xcb + (size_t)buf * DP * T;
        double* gc = gcb + buf * T * (RMAX + 2);
        for (int e = tid - first_thread; e >= 0 && e < T * DP; e += nthreads) {
            const int n = e / DP, d = e - n * DP;
            double v = 0.0;
            if (d < D && n0 + n < io.N) {
                const double xt = __dadd_rn(__dmul_rn(io.X[(size_t)(n0 + n) * D + d], gp.in_scale[d]), gp.in_shift[d]);
                v = xt / gp.ell[d];
            }
            xc[d * T + n] = v;
        }
        for (int n = tid - first_thread; n >= 0 && n < T; n += nthreads) {
            double cv = 0.0;
#pragma unroll
            for (int ro = 0; ro < RMAX; ++ro) {
                double cm = 0.0;
                if (ro < R && n0 + n < io.N) {
                    const size_t o = (size_t)(gp.slot0 + ro) * io.mstride + n0 + n;
                    cm = io.cmu[o] / gp.out_scale[ro];
                    cv += -2.0 * io.cva[o] / (gp.out_scale[ro] * gp.out_scale[ro]);
                }
                gc[n * (RMAX + 2) + ro] = cm;
            }
            gc[n * (RMAX + 2) + RMAX] = cv;
            double s2 = 0.0;                              // squared norm of the scaled candidate (the same values xcb receives)
            if (n0 + n < io.N)
                for (int d = 0; d < D; ++d) {
                    const double xt = __dadd_rn(__dmul_rn(io.X[(size_t)(n0 + n) * D + d], gp.in_scale[d]), gp.in_shift[d]);
                    const double v = xt / gp.ell[d];
                    s2 = fma(v, v, s2);
                }
            gc[n * (RMAX + 2) + RMAX + 1] = s2;
        }
    };
    const int64_t ntiles = (io.N + T - 1) / T;
    if ((int64_t)blockIdx.x < ntiles) stage_tile(blockIdx.x, 0, 0, C::THREADS);
    __syncthreads();
    if (lane == 0)
        for (int b = 0; b < C::NBUF_MAX; ++b) mbar_arrive(bar_empty + 8 * b);      // every streaming buffer starts empty

    unsigned res_par = 0, full_ph = 0, empty_ph = 0;
    int gen_buf = 0, use_buf = 0;
    int par = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * T;
        const double* xc = xcb + (size_t)par * DP * T;
        const double* gc = gcb + par * T * (RMAX + 2);
        const double* __restrict__ a_tile = io.a_scratch + (size_t)tile * io.a_stride;
        for (int e = tid; e < T * (1 + D); e += C::THREADS) gsum[e] = 0.0;          // published by the first chunk's barriers

        for (int p = 0; p < geom.nchunks; ++p) {
            const int nr = pg_chunk_rows(geom, p);
            const int row0 = p * geom.rbc;
            const int kp_end = row0 + nr;
            const bool full_chunk = nr == C::RBC;
            const bool last_chunk = p + 1 == geom.nchunks;

            double acc[RBW][NB][2];
#include "contract5_mma.inc"

            // ---- column groups of A: one bulk copy per group, issued by lane 0 of warp g mod NW
            auto request_group = [&](int g) {
                const bool mine = (w == (g & (NW - 1))) && lane == 0;
                if (g < RG) {                             // resident panel (chunk 0 only): slot g is free since the tile barrier
                    if (mine) {
                        mbar_expect_tx(bar_res + 8 * g, GS * 8);
                        bulk_copy_g2s(kres_s + g * (GS * 8), a_tile + (size_t)g * GS, GS * 8, bar_res + 8 * g);
                    }
                } else {                                  // streaming ring: wait until every warp has released the buffer
                    if (mine) {
                        mbar_wait_par(bar_empty + 8 * gen_buf, (empty_ph >> gen_buf) & 1u);
                        mbar_expect_tx(bar_full + 8 * gen_buf, GS * 8);
                        bulk_copy_g2s(kstr_s + gen_buf * (GS * 8), a_tile + (size_t)g * GS, GS * 8, bar_full + 8 * gen_buf);
                    }
                    empty_ph ^= 1u << gen_buf;            // every thread tracks the phase: the next requester may be another warp
                    gen_buf = (gen_buf + 1 == nbuf) ? 0 : gen_buf + 1;
                }
            };

            const int ngroups = (kp_end + GK - 1) / GK;
            const int g_first = p == 0 ? 0 : RG;
#pragma unroll
            for (int a = 0; a < KA; ++a) {
#pragma unroll
                for (int qd = 0; qd < QPC; ++qd) { issue_quad(qd); ring_step(); }
                advance_next();
            }
            // chunk 0: every group lands in its own panel slot, so all of them are requested at once (128 KB in flight; no
            // copy latency behind the short column steps at the tip of the triangle); later chunks: `la` groups ahead
            const int g_ahead = p == 0 ? (ngroups < RG ? ngroups : RG) : la;
            for (int g = g_first; g < g_ahead && g < ngroups; ++g) request_group(g);
#pragma unroll
            for (int r = 0; r < RBW; ++r)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) { acc[r][nb][0] = 0.0; acc[r][nb][1] = 0.0; }

            for (int g = 0; g < ngroups; ++g) {
                const int gg = g + la;
                if (gg >= g_first && gg >= g_ahead && gg < ngroups) request_group(gg);
                const double* kb;
                if (g < RG) {
                    if (p == 0) mbar_wait_par(bar_res + 8 * g, res_par);
                    kb = kres + (size_t)g * GS + lane;
                } else {
                    mbar_wait_par(bar_full + 8 * use_buf, (full_ph >> use_buf) & 1u);
                    full_ph ^= 1u << use_buf;
                    kb = kstr + (size_t)use_buf * GS + lane;
                }
                const int k0 = g * GK;
                const int k1 = (k0 + GK < kp_end) ? k0 + GK : kp_end;
                if (full_chunk && k1 + KA <= row0) {
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp)
                        mma_static(kb + (kp - k0) * (2 * NB * 32), std::integral_constant<int, 0>(), std::true_type());
                } else if (full_chunk) {
                    mma_diag_run(kb, k0, k1);         // runs of constant first row block: dispatch outside the column loop
                } else {
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp) mma_ragged(kb + (kp - k0) * (2 * NB * 32), kp);
                }
                if (g >= RG) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_empty + 8 * use_buf);
                    use_buf = (use_buf + 1 == nbuf) ? 0 : use_buf + 1;
                }
            }
            cp_async_wait<0>();

            // ---- chunk epilogue.  acc[r][nb][c] = W[i'][n], i' = 8 (row0 + r NW + w) + lane / 4, n = 8 nb + 2 (lane % 4) + c;
            // training point gi = rev - i'.  One pass per candidate block nb: the warp parks that half of its accumulators in its
            // own (drained) ring region and walks its row blocks in a compact loop -- no 16-fold unrolled body, no spills.
            double* const wreg = ring + (size_t)w * C::RING_DOUBLES_PER_WARP;      // this warp's 4 KB: accumulators, then partial tiles
            // shuffle sources that turn an accumulator-layout tile into B fragments: lane l' needs element
            // (row 4h + l' % 4, column l' / 4), held by lane 4 (4h + l' % 4) + (l' / 4) / 2 in register c = (l' / 4) % 2
            const int src0 = 4 * (lane & 3) + (lane >> 3), src1 = src0 + 16;
            const bool odd = ((lane >> 2) & 1) != 0;
            const int nrw = nr > w ? (nr - w + NW - 1) / NW : 0;                   // row blocks of this warp in the chunk
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                static_assert(RBW * 2 * 32 <= C::RING_DOUBLES_PER_WARP, "a candidate block's accumulators fit the warp's ring region");
#pragma unroll
                for (int r = 0; r < RBW; ++r)
                    *reinterpret_cast<double2*>(wreg + (r * 32 + lane) * 2) = make_double2(acc[r][nb][0], acc[r][nb][1]);
                __syncwarp();
                const double* gn0 = gc + (8 * nb + 2 * (lane & 3)) * (RMAX + 2);            // this lane's two candidates
                const double cv0 = gn0[RMAX], cv1 = gn0[RMAX + 2 + RMAX], x20 = gn0[RMAX + 1], x21 = gn0[RMAX + 2 + RMAX + 1];
                const double* __restrict__ xcol = xc + (lane & 3) * T + 8 * nb + (lane >> 2);
                double gt[4][2];                          // G^T partial tiles [column block of Xa][c] (ND <= 4)
#pragma unroll
                for (int q = 0; q < 4; ++q) { gt[q][0] = 0.0; gt[q][1] = 0.0; }
#pragma unroll 4
                for (int r = 0; r < nrw; ++r) {
                    const int gs = 8 * (row0 + r * NW + w);                        // first stream row of the block
                    const int gi = rev - (gs + (lane >> 2));                       // accumulator-layout row of this lane
                    // dot products Xs_gi . xc_n on the tensor pipe, in the accumulator layout
                    double d0 = 0.0, d1 = 0.0;
                    {
                        const double* __restrict__ xrow = Xs + (size_t)gi * D + (lane & 3);
                        for (int k = 0; k < DP; k += 4) {
                            const double a = (k + (lane & 3) < D) ? __ldg(xrow + k) : 0.0;
                            dmma884(d0, d1, a, xcol[k * T]);
                        }
                    }
                    const double xnj = __ldg(xn + gi);
                    const double2 av = *reinterpret_cast<const double2*>(wreg + (r * 32 + lane) * 2);
                    double u0 = cv0 * av.x, u1 = cv1 * av.y;
                    for (int ro = 0; ro < R; ++ro) {
                        const double al = __ldg(gp.alpha + (size_t)gi * R + ro);
                        u0 = fma(gn0[ro], al, u0);
                        u1 = fma(gn0[RMAX + 2 + ro], al, u1);
                    }
                    u0 *= dk_dr2_branchless<KIND>(variance, (-2.0 * d0 + xnj) + x20, tab32);
                    u1 *= dk_dr2_branchless<KIND>(variance, (-2.0 * d1 + xnj) + x21, tab32);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int src = h ? src1 : src0;
                        const double v0 = __shfl_sync(0xffffffffu, u0, src), v1 = __shfl_sync(0xffffffffu, u1, src);
                        const double b = odd ? v1 : v0;
                        // A fragments of the G^T product: Xa^T[c'][i], c' = 8 q + lane / 4, i = 4 h + lane % 4
                        const double* __restrict__ xk = Xs + (size_t)(rev - (gs + 4 * h + (lane & 3))) * D;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (q < ND) {                 // warp-uniform
                                const int cq = 8 * q + (lane >> 2);
                                const double xa = cq == 0 ? 1.0 : ((cq <= D) ? __ldg(xk + cq - 1) : 0.0);
                                dmma884(gt[q][0], gt[q][1], xa, b);
                            }
                        }
                    }
                }
                __syncwarp();                             // every lane has read its accumulators back: the region is free
                // gt[q] holds G^T[c' = 8 q + lane / 4][n = 8 nb + 2 (lane % 4) + c]
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (q < ND) {
                        double* o = wreg + q * 64 + (lane >> 2) * 8 + 2 * (lane & 3);
                        o[0] = gt[q][0]; o[1] = gt[q][1];
                    }
                __syncthreads();                          // (first pass: also every warp past its column loop)
                for (int e = tid; e < 8 * (1 + D); e += C::THREADS) {              // thread (n, c') of this candidate block: fixed warp order
                    const int nl = e / (1 + D), cq = e - nl * (1 + D);
                    const int off = (cq >> 3) * 64 + (cq & 7) * 8 + nl;
                    const int ge = (8 * nb + nl) * (1 + D) + cq;
                    double a = gsum[ge];
                    for (int ww = 0; ww < NW; ++ww) a += ring[(size_t)ww * C::RING_DOUBLES_PER_WARP + off];
                    gsum[ge] = a;
                }
                if (nb == NB - 1 && last_chunk && tile + gridDim.x < ntiles && tid >= C::THREADS / 2)
                    stage_tile(tile + gridDim.x, par ^ 1, C::THREADS / 2, C::THREADS / 2);
                __syncthreads();
            }
        }
        // d acq / d x_d = 2 a_d / l_d * (xs_d * G[n][0] - G[n][1 + d])
        for (int e = tid; e < T * D; e += C::THREADS) {
            const int n = e / D, d = e - n * D;
            if (n0 + n < io.N) {
                const double gval = 2.0 * gp.in_scale[d] / gp.ell[d] * (xc[d * T + n] * gsum[n * (1 + D)] - gsum[n * (1 + D) + 1 + d]);
                double* dst = io.grads + (size_t)(n0 + n) * D + d;
                *dst = io.grad_accumulate ? *dst + gval : gval;
            }
        }
        __syncthreads();                                  // gsum and xcb[par] are read above, rewritten by the next tile
        res_par ^= 1u;
        par ^= 1;
    }
}

template <int RBW, int NB, int QD, int KIND>
static inline cudaError_t launch5_bwd_kind(cudaStream_t st, const GpDev* d_gp, EvalIO io, int grid, int D) {
    using B = Contract5BwdCfg<RBW, NB, QD>;
    int nbuf = Contract5Cfg<RBW, NB, QD>::NBUF_MAX;
    while (nbuf >= 2 && B::smem_bytes(D, nbuf) > (size_t)227 * 1024) --nbuf;
    if (nbuf < 2 || !B::fits(D)) return cudaErrorInvalidConfiguration;
    const size_t smem = B::smem_bytes(D, nbuf);
    auto kern = contract5_bwd_kernel<RBW, NB, QD, KIND>;
    static size_t smem_set[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || smem_set[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 16) smem_set[dev] = smem;
    }
    kern<<<grid, 512, smem, st>>>(d_gp, io, nbuf);
    return cudaGetLastError();
}
