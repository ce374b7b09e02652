// Contraction kernel v2 (the default hot kernel): same math and work decomposition as contract.cu, re-plumbed
// after the first ncu capture (profiles/r01_contract_v1_summary.md):
//   * L^-1 blocks travel global -> shared with cp.async (LDGSTS) into a per-warp ring, S column blocks deep.
//     Every lane copies and later reads back exactly its own 16 bytes, so the ring is a private asynchronous
//     FIFO: no registers are held by loads in flight, no cross-lane synchronisation is needed, and
//     (S-1) * RBW * 512 B per warp are in flight (vs. one half column in v1, which left 23% long-scoreboard stalls).
//   * K(X*, X) generation is branch-free (covariance family is a template parameter; sqrt and exp are inlined
//     Newton / polynomial sequences without special-case paths) and processes GEN_ITERS elements per thread in
//     lock step, so the FP64 pipe sees independent chains; half of the warps of every scheduler generate the
//     next group BEFORE their DMMA work and the other half AFTER it, so generation overlaps tensor work.
//   * 16 warps x <=128 registers (4 warps per scheduler) instead of 8 x 255.
// (kernel template and launch helpers; instantiated by contract2.cu, contract2_aux.cu and contract2_prof.cu so that the three
// translation units compile in parallel)
#pragma once
#include <type_traits>
#include "contract_common.cuh"

// ABL bit 0/1: profiling ablations.  ABL bit 2 (value 4) = GRADIENT PASS: the stream is the dense block stream of
// K^-1, the accumulators hold W = K^-1 Kx, and the chunk epilogue forms u = (c_mu . alpha_i + c_var w_in) dk_in/dr^2
// and reduces u [1, Xs_i] over the training points; the tile epilogue writes d acq / d Xcand.
template <int NW, int RBW, int NB, int S, int KIND, int ABL>
__global__ void __launch_bounds__(32 * NW, 1)
contract2_kernel(const GpDev* __restrict__ gpp, const ProgramDev* __restrict__ prog, EvalIO io) {
    using C = Contract2Cfg<NW, RBW, NB, S>;
    constexpr int T = C::T;
    extern __shared__ __align__(16) double smem[];
    const int D = gpp->D;
    double* ring = smem;                                             // [NW][S][RBW][64]
    // ABL bit 2048 = PANEL (latency form for the row-split launches of a handful of candidates): all column groups of a
    // panel of up to 32 groups (2048 training points) are generated first, into 32 group buffers, then contracted without a
    // barrier per group -- a call is then a few barriers long instead of one generate -> DMMA -> barrier round per group
    constexpr bool PANEL = (ABL & 2048) != 0;
    constexpr int PANEL_GROUPS = 32;
    constexpr int KXB = PANEL ? PANEL_GROUPS : ((ABL & (8 | 32)) ? 3 : 2);   // K(X*,X) group buffers (3 in scratch mode)
    double* kx = ring + (size_t)NW * C::RING_DOUBLES_PER_WARP;       // [KXB][GROUP_SLOTS]
    double* xc = kx + KXB * C::GROUP_SLOTS;                          // [D][T]
    double* x2 = xc + (size_t)D * T;                                 // [T]
    double* tot = x2 + T;                                            // [T][1+RMAX] running colsum(A^2), A^T V
    constexpr int EW = (T + 31) / 32;                                // epilogue warps (32 candidates each)
    double* sbest = tot + T * (1 + RMAX);                            // [EW][2] best value, best index (as bits)
    double* tab32 = sbest + 2 * EW;                                  // [32] 2^(i/32)
    static_assert(C::RING_DOUBLES_PER_WARP >= T * (1 + RMAX), "the per-warp reduction slice lives in the drained ring");
    // Gradient pass, two forms.  DENSEM (bit 4): dense stream of K^-1, K(X*,X) generated here, 2 M^2 flop per candidate.
    // BWD (bit 32): the forward launch ran with ASTORE (bit 128) and parked A = L^-1 Kx per tile; this launch contracts it
    // with the reversed transposed triangular stream (row i' = Mp-1-i): W = L^-T A at M^2 flop, nothing generated.
    constexpr bool DENSEM = (ABL & 4) != 0;
    constexpr bool BWD = (ABL & 32) != 0;
    constexpr bool ASTORE = (ABL & 128) != 0;
    constexpr bool GRADM = DENSEM || BWD;
    // ABL bit 4096 = RAWP (latency path, with DENSEM + SPLIT): the launch does not wait for the program's partials
    // d acq / d (mean, var) -- it runs next to the forward launch -- and leaves the (1 + R) sums
    // sum_i w_in dk_in [1, Xs_i]  and  sum_i alpha_i,ro dk_in [1, Xs_i]  apart; thin_final_kernel combines them linearly.
    constexpr bool RAWP = (ABL & 4096) != 0;
    constexpr int GT = RAWP ? T * (1 + RMAX) : T;                    // rows of the u buffer columns / of gsum
    double* gcm = tab32 + 34;                                        // [T][RMAX]  (d acq/d mean_ro) / out_scale_ro
    double* gcv = gcm + T * RMAX;                                    // [T]        -2 sum_ro (d acq/d var_ro) / out_scale_ro^2
    double* gsum = gcv + T;                                          // [GT][1+D]  sum_i u_in [1, Xs_i]
    // ABL bit 1024 = BULK: groups re-read from the scratch arrive by one 1-D bulk copy (cp.async.bulk, the TMA engine;
    // SASS UBLKCP) issued by thread 0 and signalled on an mbarrier per group buffer, instead of four LDGSTS per thread
    constexpr bool BULK = (ABL & 1024) != 0;
    unsigned long long* kxbar = reinterpret_cast<unsigned long long*>(GRADM ? gsum + GT * (1 + D) : gcm);   // [KXB]
    static_assert(!GRADM || NW * C::RING_DOUBLES_PER_WARP + KXB * C::GROUP_SLOTS >= 8 * C::RBC * (GT + 4),
                  "the u buffer of the gradient pass reuses the ring and the group buffers");
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    const GpDev& gp = *gpp;
    const int M = gp.M, R = gp.R;
    const double variance = gp.variance;
    const double* __restrict__ Xs = gp.Xs;
    const double* __restrict__ xn = gp.xn;
    const double* __restrict__ Vv = gp.V;
    // ABL bit 16 = ROW SPLIT (few candidates): CTA (tile, split) takes chunks split, split + nsplit, ... of the block stream
    // and leaves its partial sums in io.partials; split_final_*_kernel adds them in a fixed order and runs the epilogue.
    // The stream layout may differ from the GP's default one (smaller chunks), so it travels in the launch record.
    constexpr bool SPLIT = (ABL & 16) != 0;
    const PackGeom geom = (SPLIT || BWD) ? io.geom_o : (DENSEM ? gp.geom_kinv : gp.geom);
    const double* __restrict__ packed = (SPLIT || BWD) ? io.packed_o : (DENSEM ? gp.packed_kinv : gp.packed);
    const int nsplit = SPLIT ? io.nsplit : 1;
    const int split = SPLIT ? (int)(blockIdx.x % nsplit) : 0;
    double* myring = ring + (size_t)w * C::RING_DOUBLES_PER_WARP + 2 * lane;
    const unsigned myring_s = (unsigned)__cvta_generic_to_shared(myring);     // same location, shared-window address

    const int gen_nb = w % NB;
    const int gen_n = 8 * gen_nb + (lane >> 2);
    const int gen_jl = lane & 3;
    // warps w, w+4, w+8, w+12 share a scheduler: 2 generate before their DMMA work, 2 after it.  ABL bit 256 (experiment):
    // every warp generates first, i.e. generation runs as a burst with the DMMA pipe idle
    const bool gen_first = (ABL & 256) ? true : ((w >> 2) & 1) == 0;

    constexpr bool TIM = (ABL & 64) != 0;             // profiling build: per-warp cycle counters per phase
    long long t_gen = 0, t_mma = 0, t_bar = 0, t_red = 0, t_all = 0, t_mark = 0;
    if (TIM) t_all = clock64();
    if (tid < EW) { sbest[2 * tid] = 0.0; reinterpret_cast<long long*>(sbest)[2 * tid + 1] = -1; }
    if (tid < 32) tab32[tid] = exp2(tid * (1.0 / 32.0));
    const unsigned kxbar_s = (unsigned)__cvta_generic_to_shared(kxbar);
    unsigned kx_phase = 0;                               // bit b: parity of the next completion of group buffer b's mbarrier
    if (BULK && tid == 0)
        for (int b2 = 0; b2 < KXB; ++b2) mbar_init(kxbar_s + 8 * b2, 1);
    __syncthreads();

    const int64_t ntiles = (io.N + T - 1) / T;
    for (int64_t tile = blockIdx.x / nsplit; tile < ntiles; tile += gridDim.x / nsplit) {
        const int64_t n0 = tile * T;
        __syncthreads();
        if (tid < T * (1 + RMAX)) tot[tid] = 0.0;
        if (GRADM) {
            for (int e = tid; e < GT * (1 + D); e += C::THREADS) gsum[e] = 0.0;
            if (!RAWP && tid < T) {
                const int64_t n = n0 + tid;
                double cv = 0.0;
#pragma unroll
                for (int ro = 0; ro < RMAX; ++ro) {
                    double cm = 0.0;
                    if (ro < R && n < io.N) {
                        const size_t o = (size_t)(gp.slot0 + ro) * io.mstride + n;
                        cm = io.cmu[o] / gp.out_scale[ro];
                        cv += -2.0 * io.cva[o] / (gp.out_scale[ro] * gp.out_scale[ro]);
                    }
                    gcm[tid * RMAX + ro] = cm;
                }
                gcv[tid] = cv;
            }
        }
        for (int e = tid; e < T * D; e += C::THREADS) {
            const int n = e / D, d = e - n * D;
            double x = 0.0;
            if (n0 + n < io.N) {
                if (io.gen_A) {     // RandomDesign on the device: unit-cube draw, then generative_domain >> domain (design.py:55-70, 93-94)
                    const double u = gpfo_uniform(io.gen_seed, (unsigned long long)(io.gen_first + n0 + n) * D + d);
                    x = __dadd_rn(__dmul_rn(u, io.gen_A[d]), io.gen_b[d]);
                } else {
                    x = io.X[(size_t)(n0 + n) * D + d];
                }
            }
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

        // With a K(X*,X) scratch every group is generated ONCE per tile -- by the chunk whose diagonal it belongs to, where
        // the triangular DMMA work is light and leaves pipe slots for the FP64 generation -- and parked in the per-CTA
        // scratch; later chunks copy their off-diagonal groups back asynchronously, two groups ahead of use.
        constexpr bool SCR = (ABL & (8 | 32)) != 0;      // compile-time: the plain variants carry no scratch code
        double* const scratch = BWD ? io.a_scratch + (size_t)tile * io.a_stride
                                    : (SCR ? io.kx_scratch + (size_t)blockIdx.x * io.kx_scratch_stride : nullptr);
        for (int pi = split; pi < geom.nchunks; pi += nsplit) {
            const int p = pi;
            const int nr = pg_chunk_rows(geom, p);
            const int row0 = p * geom.rbc;                // first row block of this chunk
            const int kp_end = DENSEM ? geom.nrb : row0 + nr;
            const int kd0 = DENSEM ? kp_end : row0;       // first diagonal column block (none in the dense stream)
            const double* __restrict__ cbase = packed + pg_chunk_base(geom, p) * 64 + 2 * lane;

            double acc[RBW][NB][2];
#pragma unroll
            for (int r = 0; r < RBW; ++r)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) { acc[r][nb][0] = 0.0; acc[r][nb][1] = 0.0; }

            // ---- L^-1 stream state of this warp: running source pointer (block (kp_issue, row block w)), ring cursors.
            // Block (kp, r) lives at P(kp) + r*NW*64 doubles; P advances by nr blocks per dense column block and by
            // (nr - kd - 1) per diagonal one (pg_col_base in closed form, without 64-bit multiplies in the loop).
            const double* __restrict__ src_issue = cbase + (int64_t)w * 64;
            int kp_issue = 0, wr_stage = 0, rd_stage = 0;
            // asynchronous copy of this warp's blocks of column block kp_issue into ring stage wr_stage (one commit group)
            auto issue = [&]() {
                if (!(ABL & 2) && kp_issue < kp_end) {
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
                wr_stage = (wr_stage == (S - 1) * RBW * 64) ? 0 : wr_stage + RBW * 64;
            };

            auto gen_group = [&](int g, double* __restrict__ dst) {
                double dot[C::GEN_ITERS];
                int jj[C::GEN_ITERS];
#pragma unroll
                for (int q = 0; q < C::GEN_ITERS; ++q) {
                    const int ksl = (w + NW * q) / NB;
                    const int j = (g * KPC) * 8 + 4 * ksl + gen_jl;
                    jj[q] = j < M ? j : M - 1;            // clamp: keeps the loads in range, result masked below
                    dot[q] = 0.0;
                }
                if (!(ABL & 1)) {
                    for (int d = 0; d < D; ++d) {
                        const double c = xc[d * T + gen_n];
#pragma unroll
                        for (int q = 0; q < C::GEN_ITERS; ++q) dot[q] = fma(__ldg(Xs + (size_t)jj[q] * D + d), c, dot[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < C::GEN_ITERS; ++q) {
                    const int ksl = (w + NW * q) / NB;
                    const int j = (g * KPC) * 8 + 4 * ksl + gen_jl;
                    double val;
                    if (ABL & 1) val = 1e-3 * j;
                    else {
                        const double r2 = (-2.0 * dot[q] + __ldg(xn + jj[q])) + x2[gen_n];    // square_dist, expanded form
                        val = kx_branchless<KIND>(variance, r2, tab32);
                    }
                    const double out = j < M ? val : 0.0;
                    dst[(ksl * NB + gen_nb) * 32 + lane] = out;
                    if (SCR) __stcg(scratch + (size_t)g * C::GROUP_SLOTS + (ksl * NB + gen_nb) * 32 + lane, out);
                }
                // the scratch is re-read through the async proxy: order this thread's generic writes before it
                if (BULK) asm volatile("fence.proxy.async;" ::: "memory");
            };

            // scratch mode, re-read chunks: asynchronous copy of a whole group scratch -> shared, issued TWO groups ahead
            // (third buffer), so its L2/HBM latency never sits between the DMMA work and the group barrier
            auto copy_group = [&](int g) {
                const double* __restrict__ src = scratch + (size_t)g * C::GROUP_SLOTS;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(kx + (g % KXB) * C::GROUP_SLOTS);
                if (BULK) {
                    if (tid == 0) {
                        const unsigned bar = kxbar_s + 8 * (g % KXB);
                        constexpr unsigned bytes = C::GROUP_SLOTS * sizeof(double);
                        asm volatile("fence.proxy.async;" ::: "memory");
                        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}"
                                     ::"r"(bar), "r"(bytes) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < C::GROUP_SLOTS / (2 * C::THREADS); ++q) {
                        const int ch = tid + q * C::THREADS;             // 16-byte chunk index
                        cp_async16(dst + ch * 16, src + 2 * ch);
                    }
                    cp_async_commit();
                }
            };

            // one column block: B fragments from the Kx group buffer, A fragments from the ring, RBW*2*NB DMMAs.
            // FULL = every row block of this warp is active (dense column block of a full chunk): no predicates.
            auto mma_kp = [&](const double* __restrict__ kbk, int lo, auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                issue();
                double bf[2][NB];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) bf[h][nb] = kbk[(h * NB + nb) * 32];
                cp_async_wait<S - 1>();               // this thread's copies for the current column block have landed
                const double* stage = myring + rd_stage;
#pragma unroll
                for (int r = 0; r < RBW; ++r) {
                    const int rbl = r * NW + w;
                    if (FULL || (rbl >= lo && rbl < nr)) {
                        double2 a;
                        if (ABL & 2) a = make_double2(1e-3 * lane, 2e-3);
                        else a = *reinterpret_cast<const double2*>(stage + r * 64);
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) dmma884(acc[r][nb][0], acc[r][nb][1], a.x, bf[0][nb]);
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) dmma884(acc[r][nb][0], acc[r][nb][1], a.y, bf[1][nb]);
                    }
                }
                rd_stage = (rd_stage == (S - 1) * RBW * 64) ? 0 : rd_stage + RBW * 64;
            };
            auto mma_group = [&](int g) {
                const double* __restrict__ kb = kx + (g % KXB) * C::GROUP_SLOTS + lane;      // PANEL: KXB = groups per panel
                const int k0 = g * KPC;
                const int k1 = (k0 + KPC < kp_end) ? k0 + KPC : kp_end;
                if (nr == C::RBC && k1 <= kd0) {       // whole group in the dense part of a full chunk
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp) mma_kp(kb + (kp - k0) * (2 * NB * 32), 0, std::true_type());
                } else {
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp)
                        mma_kp(kb + (kp - k0) * (2 * NB * 32), kp > kd0 ? kp - kd0 : 0, std::false_type());
                }
            };

            const int ngroups = (kp_end + KPC - 1) / KPC;
            // groups left of the diagonal were generated by earlier chunks (backward pass: every group is a stored one)
            const int nold = BWD ? ngroups : (SCR ? kd0 / KPC : 0);
            if (PANEL) {
#pragma unroll
                for (int s = 0; s < S - 1; ++s) issue();
                for (int g0 = 0; g0 < ngroups; g0 += PANEL_GROUPS) {
                    const int g1 = g0 + PANEL_GROUPS < ngroups ? g0 + PANEL_GROUPS : ngroups;
                    __syncthreads();                                  // the previous panel (or chunk) has been consumed
#pragma unroll 4
                    for (int g = g0; g < g1; ++g) gen_group(g, kx + (g % KXB) * C::GROUP_SLOTS);
                    __syncthreads();
                    for (int g = g0; g < g1; ++g) mma_group(g);
                }
                __syncthreads();
            } else {
            if (nold > 0) {
                copy_group(0);
                if (nold > 1) copy_group(1);
                if (!BULK) cp_async_wait<0>();
            } else if (!BWD) {
                gen_group(0, kx);
            }
#pragma unroll
            for (int s = 0; s < S - 1; ++s) issue();
            __syncthreads();
            for (int g = 0; g < ngroups; ++g) {
                const bool more = g + 1 < ngroups;
                double* nxt = kx + ((g + 1) % KXB) * C::GROUP_SLOTS;
                // (A split arrive/wait barrier on an mbarrier was measured here: it moves waiting time from the barrier
                // into pipe contention, 26.8 vs 27.2 TFLOP/s, so the plain CTA barrier stays.)
                if (SCR && g + 2 < nold) copy_group(g + 2);
                if (BULK && g < nold) {               // this group was bulk-copied: wait for its bytes
                    unsigned par = (kx_phase >> (g % KXB)) & 1u;
                    mbar_wait(kxbar_s + 8 * (g % KXB), par);
                    kx_phase ^= 1u << (g % KXB);
                }
                if (SCR && g + 1 < nold) {            // next group comes from the scratch (already in flight)
                    if (TIM) t_mark = clock64();
                    mma_group(g);
                    if (TIM) t_mma += clock64() - t_mark;
                    // it was committed a whole iteration ago; only a very short group can leave it pending
                    const int kcount = (kp_end < (g + 1) * KPC ? kp_end : (g + 1) * KPC) - g * KPC;
                    if (!BULK && kcount < S) cp_async_wait<0>();
                } else if (gen_first) {
                    if (TIM) t_mark = clock64();
                    if (!BWD && more) gen_group(g + 1, nxt);
                    if (TIM) { const long long t = clock64(); t_gen += t - t_mark; t_mark = t; }
                    mma_group(g);
                    if (TIM) t_mma += clock64() - t_mark;
                } else {
                    if (TIM) t_mark = clock64();
                    mma_group(g);
                    if (TIM) { const long long t = clock64(); t_mma += t - t_mark; t_mark = t; }
                    if (!BWD && more) gen_group(g + 1, nxt);
                    if (TIM) t_gen += clock64() - t_mark;
                }
                if (TIM) t_mark = clock64();
                __syncthreads();
                if (TIM) {                            // BAR.SYNC defers blocking to the next dependent access: force one
                    if (*reinterpret_cast<volatile double*>(tab32) == 123.456) t_gen += 1;
                    t_bar += clock64() - t_mark;
                }
            }
            }   // !PANEL
            cp_async_wait<0>();

            if (GRADM) {
                // ---- gradient pass: u_in = (sum_ro c_mu[n][ro] alpha[i][ro] + c_var[n] w_in) * dk_in/dr^2 into the (drained) ring
                __syncthreads();                          // every warp is done reading its ring
                constexpr int US = GT + 4;                // row stride of the u buffer: 4 mod 16 doubles, so that the A-fragment reads
                                                          // of the reduction below (4 rows x 4 candidates per half-warp) hit 16 banks
                double* ubuf = ring;                      // [8 * RBC][US]  (spills into the group buffers when S < NB)
                const int rev = 8 * geom.nrb - 1;         // backward pass: stream row i' is training point rev - i'
#pragma unroll
                for (int r = 0; r < RBW; ++r) {
                    const int rbl = r * NW + w;
                    const int il = 8 * rbl + (lane >> 2);
                    const int gs = 8 * (row0 + rbl) + (lane >> 2);
                    const int gi = BWD ? rev - gs : gs;
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const int n = 8 * nb + 2 * (lane & 3) + c;
                            double u = 0.0;
                            if (RAWP) {
                                double dk = 0.0;
                                if (rbl < nr) {
                                    double dot = 0.0;
                                    for (int d = 0; d < D; ++d) dot = fma(__ldg(Xs + (size_t)gi * D + d), xc[d * T + n], dot);
                                    const double r2 = (-2.0 * dot + __ldg(xn + gi)) + x2[n];
                                    dk = dk_dr2_branchless<KIND>(variance, r2, tab32);
                                }
                                ubuf[il * US + n] = acc[r][nb][c] * dk;
                                for (int ro = 0; ro < R; ++ro)
                                    ubuf[il * US + T * (1 + ro) + n] = rbl < nr ? __ldg(gp.alpha + (size_t)gi * R + ro) * dk : 0.0;
                                continue;
                            }
                            if (rbl < nr) {
                                double coef = gcv[n] * acc[r][nb][c];
                                for (int ro = 0; ro < R; ++ro) coef = fma(gcm[n * RMAX + ro], __ldg(gp.alpha + (size_t)gi * R + ro), coef);
                                // (regenerating dk/dr^2 here instead of storing it in the forward pass costs ~1 ms of a 43 ms
                                // backward launch at M = 2048: measured by ablation, not worth a second scratch)
                                double dot = 0.0;
                                for (int d = 0; d < D; ++d) dot = fma(__ldg(Xs + (size_t)gi * D + d), xc[d * T + n], dot);
                                const double r2 = (-2.0 * dot + __ldg(xn + gi)) + x2[n];
                                u = coef * dk_dr2_branchless<KIND>(variance, r2, tab32);
                            }
                            ubuf[il * US + n] = u;
                        }
                }
                __syncthreads();
                // ---- G[n][c'] += sum_i u[i][n] * Xa[i][c'],  Xa = [1, Xs]: a (T x rows) x (rows x (1+D)) product, done as 8 x 8
                // tiles of DMMAs (one or a few tiles per warp: candidate block x column block) instead of T (1+D) scalar row
                // loops -- with T (1+D) = 576 tasks on 512 threads the scalar form ran a second round on 64 threads.
                // Two interleaved accumulator pairs (even / odd k-steps) keep the dependent DMMA chain short; row order fixed.
                {
                    const int ND = (1 + D + 7) / 8;       // column blocks of Xa
                    const int rows = 8 * nr;              // a multiple of 8
                    const int step = BWD ? -1 : 1;
                    const int g0 = BWD ? rev - 8 * row0 : 8 * row0;             // training point of buffer row 0
                    const int NBQ = RAWP ? NB * (1 + R) : NB;     // 8-wide column blocks of the u buffer
                    for (int q = w; q < NBQ * ND; q += NW) {
                        const int nbq = q % NBQ, dbq = q / NBQ;
                        const int cq = 8 * dbq + (lane >> 2);                   // column of Xa this lane feeds (B fragment: col = lane / 4)
                        const double* __restrict__ up = ubuf + (lane & 3) * US + 8 * nbq + (lane >> 2);
                        const bool is_one = cq == 0, is_x = cq >= 1 && cq <= D;
                        const double* __restrict__ xp = Xs + (size_t)(g0 + step * (lane & 3)) * D + (is_x ? cq - 1 : 0);
                        double e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0;
                        for (int k = 0; k < rows; k += 8) {
                            const double a0 = up[k * US], a1 = up[(k + 4) * US];
                            const double b0 = is_one ? 1.0 : (is_x ? __ldg(xp + (ptrdiff_t)step * k * D) : 0.0);
                            const double b1 = is_one ? 1.0 : (is_x ? __ldg(xp + (ptrdiff_t)step * (k + 4) * D) : 0.0);
                            dmma884(e0, e1, a0, b0);
                            dmma884(o0, o1, a1, b1);
                        }
                        // lane holds G[n = 8 nbq + lane / 4][c' = 8 dbq + 2 (lane % 4) + {0, 1}]; every element has one owner
                        const int n = 8 * nbq + (lane >> 2), c0 = 8 * dbq + 2 * (lane & 3);
                        if (c0 <= D) gsum[n * (1 + D) + c0] += e0 + o0;
                        if (c0 + 1 <= D) gsum[n * (1 + D) + c0 + 1] += e1 + o1;
                    }
                }
                __syncthreads();
                continue;
            }
            if (ASTORE) {
                // park A of this chunk for the backward launch: reversed row order, B-fragment group layout (the slot of
                // element (j', n) is the one gen_group uses for K(X*,X) element (j', n))
                double* const at = io.a_scratch + (size_t)tile * io.a_stride;
                const int i8 = 7 - (lane >> 2);
#pragma unroll
                for (int r = 0; r < RBW; ++r) {
                    const int rbl = r * NW + w;
                    if (rbl < nr) {
                        const int rbrev = geom.nrb - 1 - (row0 + rbl);
                        double* const o = at + (size_t)(rbrev >> 3) * C::GROUP_SLOTS +
                                          (((rbrev & 7) * 2 + (i8 >> 2)) * NB) * 32 + (i8 & 3) + 8 * (lane & 3);
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) {
                            __stcg(o + nb * 32, acc[r][nb][0]);
                            __stcg(o + nb * 32 + 4, acc[r][nb][1]);
                        }
                    }
                }
            }
            if (TIM) t_mark = clock64();
            // ---- chunk reduction (fixed order -> deterministic).  Warp w's slice red[w][T][1+RMAX] lives at the start of its
            // own, now drained, ring region; the next chunk's copies are issued only after the barriers below.
            double* const red = ring;
            constexpr int RS = C::RING_DOUBLES_PER_WARP;             // warp stride of the aliased slices
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
            __syncthreads();
            if (tid < T * (1 + RMAX)) {                  // thread (n, k): fixed warp order -> deterministic
                const int n = tid / (1 + RMAX), k = tid - n * (1 + RMAX);
                if (k <= R) {
                    double a = tot[tid];
                    for (int ww = 0; ww < NW; ++ww) a += red[ww * RS + n * (1 + RMAX) + k];
                    tot[tid] = a;
                }
            }
            __syncthreads();
            if (TIM) t_red += clock64() - t_mark;
        }
        if (SPLIT) {
            const int gt = RAWP ? T * (1 + R) : T;
            double* part = io.partials + ((size_t)tile * nsplit + split) * (size_t)(GRADM ? gt * (1 + D) : T * (1 + RMAX));
            if (GRADM) for (int e = tid; e < gt * (1 + D); e += C::THREADS) part[e] = gsum[e];
            else if (tid < T * (1 + RMAX)) part[tid] = tot[tid];
            continue;
        }
        if (GRADM) {
            // d acq / d x_d = 2 a_d / l_d * (xs_d * sum_i u_i - sum_i u_i Xs_id)       (d r^2 / d x_d = 2 (xs_d - Xs_id) a_d / l_d)
            for (int e = tid; e < T * D; e += C::THREADS) {
                const int n = e / D, d = e - n * D;
                if (n0 + n < io.N) {
                    const double g = 2.0 * gp.in_scale[d] / gp.ell[d] *
                                     (xc[d * T + n] * gsum[n * (1 + D)] - gsum[n * (1 + D) + 1 + d]);
                    double* dst = io.grads + (size_t)(n0 + n) * D + d;
                    *dst = io.grad_accumulate ? *dst + g : g;
                }
            }
            continue;
        }
        if (w < EW) {                                     // epilogue warp w owns candidates 32w .. 32w+31 of the tile
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
    if (TIM && lane == 0 && io.debug) {
        long long* o = io.debug + ((size_t)blockIdx.x * NW + w) * 8;
        o[0] = clock64() - t_all; o[1] = t_gen; o[2] = t_mma; o[3] = t_bar; o[4] = t_red;
    }
    __syncthreads();
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


template <int NW, int RBW, int NB, int S, int KIND, int ABL>
static inline cudaError_t launch2_kind(cudaStream_t st, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int grid, int D) {
    using C = Contract2Cfg<NW, RBW, NB, S>;
    const size_t smem = ((size_t)NW * C::RING_DOUBLES_PER_WARP + ((ABL & 2048) ? 32 : ((ABL & (8 | 32)) ? 3 : 2)) * C::GROUP_SLOTS +
                         (size_t)D * C::T + C::T + C::T * (1 + RMAX) + 2 * ((C::T + 31) / 32) + 34 +
                         ((ABL & (4 | 32)) ? C::T * RMAX + C::T + ((ABL & 4096) ? 1 + RMAX : 1) * C::T * (1 + (size_t)D) : 0) +
                         ((ABL & 1024) ? 4 : 0)) * sizeof(double);
    auto kern = contract2_kernel<NW, RBW, NB, S, KIND, ABL>;
    static size_t smem_set[16] = {0};              // per instantiation and device: the opt-in is sticky, raise it only when needed
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || smem_set[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 16) smem_set[dev] = smem;
    }
    kern<<<grid, C::THREADS, smem, st>>>(d_gp, d_prog, io);
    return cudaGetLastError();
}

template <int NW, int RBW, int NB, int S, int ABL = 0>
static inline cudaError_t launch2(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int grid,
                           int D) {
    switch (kind) {
        case GPFO_KERN_RBF: return launch2_kind<NW, RBW, NB, S, GPFO_KERN_RBF, ABL>(st, d_gp, d_prog, io, grid, D);
        case GPFO_KERN_MATERN52: return launch2_kind<NW, RBW, NB, S, GPFO_KERN_MATERN52, ABL>(st, d_gp, d_prog, io, grid, D);
        default: return launch2_kind<NW, RBW, NB, S, GPFO_KERN_MATERN32, ABL>(st, d_gp, d_prog, io, grid, D);
    }
}

// Ring depth of the 64-candidate tiling (variant 12 and its gradient-pass relatives).  Measured at M = 2048: S = 3: 27.23,
// 4: 26.95, 5: 27.34, 6: 27.33, 7: 27.36 TF/s -- five stages of 1 KB per warp
constexpr int S12 = 5;
