// Contraction kernel v3: v2 (contract2.cu) run by a PAIR of CTAs per candidate tile (thread-block cluster of 2).
//
// Why: DFMA and DMMA share one pipe on B200 (profiles/r01_fp64_pipes.log), so the FP64 instructions that generate
// K(X*, X) are taken from the contraction.  With 512-row chunks per CTA a lone CTA regenerates the covariance columns
// for each of its chunks (2.5x at M = 2048).  Here the two CTAs of a cluster own the row blocks of opposite parity of
// 1024-row cluster chunks, walk the same column blocks in lock step, and each generates HALF of every K(X*, X) group,
// storing it into its own and (through distributed shared memory) its partner's buffer: per-CTA generation work drops
// from 2.5 M T to 0.75 M T evaluations at M = 2048.  One cluster barrier per group replaces the CTA barrier.
// colsum(A^2) and A^T V of the odd-parity CTA travel to the even one through DSMEM; rank 0 runs the epilogue.
#include <cooperative_groups.h>
#include "contract_common.cuh"

namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned map_to_peer(unsigned saddr, unsigned peer) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(peer));
    return r;
}
__device__ __forceinline__ void st_cluster_f64(unsigned caddr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(caddr), "d"(v) : "memory");
}
// split barrier: arrive as soon as this thread's stores are done (release), wait just before the first dependent load
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int NW, int RBW, int NB, int S, int KIND>
__global__ void __launch_bounds__(32 * NW, 1)
contract3_kernel(const GpDev* __restrict__ gpp, const ProgramDev* __restrict__ prog, EvalIO io) {
    using C = Contract2Cfg<NW, RBW, NB, S>;
    constexpr int T = C::T;
    constexpr int HALF_ITERS = C::GEN_ITERS / 2;                     // elements per thread per group (this CTA's half)
    static_assert(C::GEN_ITERS % 2 == 0, "the group must split evenly over the two CTAs");
    extern __shared__ __align__(16) double smem[];
    const int D = gpp->D;
    double* ring = smem;                                             // [NW][S][RBW][64]
    double* kx = ring + (size_t)NW * C::RING_DOUBLES_PER_WARP;       // [3][GROUP_SLOTS]  (triple buffer, see group loop)
    double* red = kx + 3 * C::GROUP_SLOTS;                           // [NW][T][1+RMAX]
    double* xc = red + (size_t)NW * T * (1 + RMAX);                  // [D][T]
    double* x2 = xc + (size_t)D * T;                                 // [T]
    double* tot = x2 + T;                                            // [T][1+RMAX]
    double* tot_peer = tot + T * (1 + RMAX);                         // [T][1+RMAX]  written by the partner CTA
    double* sbest = tot_peer + T * (1 + RMAX);                       // [2]
    double* tab32 = sbest + 2;                                       // [32]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    unsigned crank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int c = (int)crank;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

    const GpDev& gp = *gpp;
    const int M = gp.M, R = gp.R;
    const double variance = gp.variance;
    const double* __restrict__ Xs = gp.Xs;
    const double* __restrict__ xn = gp.xn;
    const double* __restrict__ Vv = gp.V;
    const double* __restrict__ packed = gp.packed_c[c];
    const int nrb = gp.cgeom.nrb, rbc = gp.cgeom.rbc, nchunks = gp.cgeom.nchunks;
    double* myring = ring + (size_t)w * C::RING_DOUBLES_PER_WARP + 2 * lane;
    const unsigned myring_s = (unsigned)__cvta_generic_to_shared(myring);
    const unsigned kx_peer = map_to_peer((unsigned)__cvta_generic_to_shared(kx), crank ^ 1u);
    const unsigned tot_peer_remote = map_to_peer((unsigned)__cvta_generic_to_shared(tot_peer), crank ^ 1u);

    const int gen_nb = w % NB;
    const int gen_n = 8 * gen_nb + (lane >> 2);
    const int gen_jl = lane & 3;
    const bool gen_first = ((w >> 2) & 1) == 0;

    if (tid == 0) { sbest[0] = 0.0; reinterpret_cast<long long*>(sbest)[1] = -1; }
    if (tid < 32) tab32[tid] = exp2(tid * (1.0 / 32.0));

    const int64_t ntiles = (io.N + T - 1) / T;
    for (int64_t tile = cluster_id; tile < ntiles; tile += nclusters) {
        const int64_t n0 = tile * T;
        cluster_sync_all();                               // partner is done with this CTA's tot_peer / kx buffers
        if (tid < T * (1 + RMAX)) tot[tid] = 0.0;
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

        for (int p = 0; p < nchunks; ++p) {
            const int nr = cg_rows(gp.cgeom, p, c);       // this CTA's row blocks in the cluster chunk
            const int kd0 = p * 2 * rbc;                  // first diagonal column block of the cluster chunk
            const int kp_end = (kd0 + 2 * rbc < nrb) ? kd0 + 2 * rbc : nrb;

            double acc[RBW][NB][2];
#pragma unroll
            for (int r = 0; r < RBW; ++r)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) { acc[r][nb][0] = 0.0; acc[r][nb][1] = 0.0; }

            // block (kp, q) of this rank lives at P(kp) + q*64; P advances by nr - lo(kd) - (lo(kd+1) - lo(kd)) blocks
            const double* __restrict__ src_issue = packed + gp.cgeom.base[c][p] * 64 + 2 * lane + (int64_t)w * 64;
            int kp_issue = 0, wr_stage = 0, rd_stage = 0;
            auto issue = [&]() {
                if (kp_issue < kp_end) {
                    const int kd = kp_issue - kd0;
                    const int lo = cg_lo(kd, c);
                    const unsigned dst = myring_s + wr_stage * 8;
#pragma unroll
                    for (int r = 0; r < RBW; ++r) {
                        const int q = r * NW + w;
                        if (q >= lo && q < nr) cp_async16(dst + r * 512, src_issue + r * (NW * 64));
                    }
                    src_issue += (nr - cg_lo(kd + 1, c)) * 64;
                }
                cp_async_commit();
                ++kp_issue;
                wr_stage = (wr_stage == (S - 1) * RBW * 64) ? 0 : wr_stage + RBW * 64;
            };

            // this CTA's half of group g: k-steps of its parity; every value goes to both CTAs' buffers
            auto gen_half = [&](int g, int buf) {
                double dot[HALF_ITERS];
                int jj[HALF_ITERS];
#pragma unroll
                for (int q = 0; q < HALF_ITERS; ++q) {
                    const int ksl = (w + NW * (2 * q + c)) / NB;
                    const int j = (g * KPC) * 8 + 4 * ksl + gen_jl;
                    jj[q] = j < M ? j : M - 1;
                    dot[q] = 0.0;
                }
                for (int d = 0; d < D; ++d) {
                    const double xv = xc[d * T + gen_n];
#pragma unroll
                    for (int q = 0; q < HALF_ITERS; ++q) dot[q] = fma(__ldg(Xs + (size_t)jj[q] * D + d), xv, dot[q]);
                }
#pragma unroll
                for (int q = 0; q < HALF_ITERS; ++q) {
                    const int ksl = (w + NW * (2 * q + c)) / NB;
                    const int j = (g * KPC) * 8 + 4 * ksl + gen_jl;
                    const double r2 = (-2.0 * dot[q] + __ldg(xn + jj[q])) + x2[gen_n];
                    const double val = j < M ? kx_branchless<KIND>(variance, r2, tab32) : 0.0;
                    const int slot = buf * C::GROUP_SLOTS + (ksl * NB + gen_nb) * 32 + lane;
                    kx[slot] = val;
                    st_cluster_f64(kx_peer + slot * 8, val);
                }
            };

            auto mma_kp = [&](const double* __restrict__ kbk, int lo, auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                issue();
                double bf[2][NB];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) bf[h][nb] = kbk[(h * NB + nb) * 32];
                cp_async_wait<S - 1>();
                const double* stage = myring + rd_stage;
#pragma unroll
                for (int r = 0; r < RBW; ++r) {
                    const int q = r * NW + w;
                    if (FULL || (q >= lo && q < nr)) {
                        const double2 a = *reinterpret_cast<const double2*>(stage + r * 64);
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) dmma884(acc[r][nb][0], acc[r][nb][1], a.x, bf[0][nb]);
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) dmma884(acc[r][nb][0], acc[r][nb][1], a.y, bf[1][nb]);
                    }
                }
                rd_stage = (rd_stage == (S - 1) * RBW * 64) ? 0 : rd_stage + RBW * 64;
            };
            auto mma_group = [&](int g) {
                const double* __restrict__ kb = kx + (g % 3) * C::GROUP_SLOTS + lane;
                const int k0 = g * KPC;
                const int k1 = (k0 + KPC < kp_end) ? k0 + KPC : kp_end;
                if (nr == C::RBC && k1 <= kd0 + c) {      // lo == 0 for kd <= c - 1 ... and every row block exists
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp) mma_kp(kb + (kp - k0) * (2 * NB * 32), 0, std::true_type());
                } else {
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp) mma_kp(kb + (kp - k0) * (2 * NB * 32), cg_lo(kp - kd0, c), std::false_type());
                }
            };

            const int ngroups = (kp_end + KPC - 1) / KPC;
            // Group pipeline with a split cluster barrier and three K(X*,X) buffers: a thread ARRIVES as soon as its share of
            // group g+1 is stored (release) and WAITS only before group g+1 is first read, so the barrier latency hides
            // behind the DMMA work of group g.  Buffer (g+1)%3 was last read by mma(g-2), which every thread finished before
            // its arrive of iteration g-1, i.e. before anyone passes the wait that precedes these stores (2 buffers would race).
            gen_half(0, 0);
            cluster_arrive();
#pragma unroll
            for (int s = 0; s < S - 1; ++s) issue();
            cluster_wait();
            for (int g = 0; g < ngroups; ++g) {
                const bool more = g + 1 < ngroups;
                if (gen_first) {
                    if (more) gen_half(g + 1, (g + 1) % 3);
                    cluster_arrive();
                    mma_group(g);
                } else {
                    mma_group(g);
                    if (more) gen_half(g + 1, (g + 1) % 3);
                    cluster_arrive();
                }
                cluster_wait();
            }
            cp_async_wait<0>();
            cluster_sync_all();                           // every warp of both CTAs is done reading the K(X*,X) buffers

            // ---- chunk reduction over this CTA's rows (fixed order -> deterministic)
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
                    for (int cc = 0; cc < 2; ++cc) {
                        double v = ss[nb][cc];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        if (lane < 4) red[(w * T + nb * 8 + 2 * lane + cc) * (1 + RMAX)] = v;
                    }
                for (int ro = 0; ro < R; ++ro) {
                    double mv[NB][2];
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) { mv[nb][0] = 0.0; mv[nb][1] = 0.0; }
#pragma unroll
                    for (int r = 0; r < RBW; ++r) {
                        const int q = r * NW + w;
                        double v = 0.0;
                        if (q < nr) v = __ldg(Vv + (size_t)(8 * (kd0 + 2 * q + c) + (lane >> 2)) * R + ro);
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) {
                            mv[nb][0] += acc[r][nb][0] * v;
                            mv[nb][1] += acc[r][nb][1] * v;
                        }
                    }
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) {
                            double v = mv[nb][cc];
                            v += __shfl_xor_sync(0xffffffffu, v, 4);
                            v += __shfl_xor_sync(0xffffffffu, v, 8);
                            v += __shfl_xor_sync(0xffffffffu, v, 16);
                            if (lane < 4) red[(w * T + nb * 8 + 2 * lane + cc) * (1 + RMAX) + 1 + ro] = v;
                        }
                }
            }
            __syncthreads();
            if (tid < T * (1 + RMAX)) {
                const int n = tid / (1 + RMAX), k = tid - n * (1 + RMAX);
                if (k <= R) {
                    double a = tot[tid];
                    for (int ww = 0; ww < NW; ++ww) a += red[(ww * T + n) * (1 + RMAX) + k];
                    tot[tid] = a;
                }
            }
            __syncthreads();
        }
        // ---- the odd-parity CTA hands its sums to the even one; rank 0 runs the epilogue (fixed order: even + odd)
        if (c == 1 && tid < T * (1 + RMAX)) st_cluster_f64(tot_peer_remote + tid * 8, tot[tid]);
        cluster_sync_all();
        if (c == 0 && w == 0) {
            double tot_mv[RMAX];
            const int ln = lane < T ? lane : 0;
#pragma unroll
            for (int ro = 0; ro < RMAX; ++ro) tot_mv[ro] = tot[ln * (1 + RMAX) + 1 + ro] + tot_peer[ln * (1 + RMAX) + 1 + ro];
            const double tss = tot[ln * (1 + RMAX)] + tot_peer[ln * (1 + RMAX)];
            double best_v = sbest[0];
            long long best_i = reinterpret_cast<long long*>(sbest)[1];
            gpfo_tile_epilogue<T>(gp, prog, io, n0, lane, tss, tot_mv, best_v, best_i);
            if (lane == 0) { sbest[0] = best_v; reinterpret_cast<long long*>(sbest)[1] = best_i; }
        }
    }
    cluster_sync_all();                                   // no CTA exits while its partner may still store into it
    if (c == 0 && tid == 0 && io.run_program) {
        io.part_val[cluster_id] = sbest[0];
        io.part_idx[cluster_id] = reinterpret_cast<long long*>(sbest)[1];
    }
}

template <int NW, int RBW, int NB, int S, int KIND>
static cudaError_t launch3_kind(cudaStream_t st, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int nclusters, int D) {
    using C = Contract2Cfg<NW, RBW, NB, S>;
    const size_t smem = ((size_t)NW * C::RING_DOUBLES_PER_WARP + 3 * C::GROUP_SLOTS + (size_t)NW * C::T * (1 + RMAX) +
                         (size_t)D * C::T + C::T + 2 * C::T * (1 + RMAX) + 2 + 32) * sizeof(double);
    auto kern = contract3_kernel<NW, RBW, NB, S, KIND>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * nclusters);
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, d_gp, d_prog, io);
}

cudaError_t gpfo_contract3_launch(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io,
                                  int nclusters, int D) {
    switch (kind) {
        case GPFO_KERN_RBF: return launch3_kind<16, 4, 4, 4, GPFO_KERN_RBF>(st, d_gp, d_prog, io, nclusters, D);
        case GPFO_KERN_MATERN52: return launch3_kind<16, 4, 4, 4, GPFO_KERN_MATERN52>(st, d_gp, d_prog, io, nclusters, D);
        default: return launch3_kind<16, 4, 4, 4, GPFO_KERN_MATERN32>(st, d_gp, d_prog, io, nclusters, D);
    }
}
int gpfo_contract3_rbc() { return 64; }
int gpfo_contract3_clusters(int64_t N, int num_sms) {
    const int64_t ntiles = (N + 31) / 32;
    const int cap = num_sms / 2;
    return (int)(ntiles < cap ? (ntiles > 0 ? ntiles : 1) : cap);
}
