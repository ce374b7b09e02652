// Contraction kernel v5 ("resident panel"): K(X*,X) stays on the SM -- no global scratch, no HBM traffic beyond one read of
// the candidates and of the L^-1 block stream.
//
// Why a new tiling.  The accumulators of a CTA hold (chunk rows) x (tile width) = 16 K doubles.  v2
// (contract2_kernel.cuh) spends that budget on wide tiles (T = 64, 256-row chunks), so a tile is walked in 8 row chunks and
// every chunk needs all covariance columns left of its diagonal: either regenerate them (4.5 x the M T evaluations at
// M = 2048, on the FP64 pipe the DMMAs share) or park them in global memory (the 148 MB scratch of variants 11/12, which the
// L2 cannot hold: 44 GB of DRAM traffic per 10^6 candidates).  v5 spends the budget on tall chunks instead:
//   * T = 16 candidates x 1024-row chunks (RBW = 8 row blocks per warp, NB = 2 candidate blocks), or T = 32 x 512 rows.
//   * The covariance columns of chunk-column 0 (the first 1024 / 512 training points, 128 KB) are generated ONCE per tile into
//     a RESIDENT shared-memory panel; every later row chunk re-reads them from shared memory.  At M <= 2048 (T = 16) nothing
//     is ever generated twice and nothing is parked anywhere; at M = 4096 / 8192 the columns right of the panel are
//     regenerated per chunk (1.75 x / 3.6 x, on 1/2 / 1/4 of the per-flop cost of M = 2048).
//   * Tall chunks also balance the diagonal triangles: a warp owns 8 row blocks dealt cyclically, so inside a triangle the
//     per-warp work differs by at most one block-step in eight (v2: one in two).
//   * No CTA barrier in the column loop.  Column groups (8 KB each) are handed over through mbarriers: every warp arrives on
//     full[g] after writing its share of group g (generated two groups ahead of use) and waits on it right before the first
//     B-fragment load; the streaming groups right of the panel cycle through a small ring with empty[] barriers.  Warps of a
//     CTA may drift apart by a whole group instead of meeting at a __syncthreads every 8 column blocks.
//   * L^-1 blocks travel global -> shared with cp.async into a per-warp private ring of QD quads of row blocks (each lane
//     copies and re-reads its own 16 bytes: no barrier); the slot just read is refilled for the column block one ahead.
//     One 512-byte block feeds 2 NB DMMAs, so the stream is 4 x (T = 16) the L2 traffic of v2's T = 64 -- 7 TB/s at
//     M = 2048, which the L2 (about 12 TB/s) carries; that is the price of keeping K(X*,X) on chip.
#pragma once
#include <type_traits>
#include "contract_common.cuh"

template <int RBW, int NB, int QD>
struct Contract5Cfg {
    static constexpr int NW = 16;
    static constexpr int THREADS = 32 * NW;
    static constexpr int T = 8 * NB;                        // candidates per tile
    static constexpr int RBC = NW * RBW;                    // row blocks per chunk
    static constexpr int GK = 16 / NB;                      // column blocks per K(X*,X) group: every group is 8 KB
    static constexpr int GROUP_SLOTS = GK * 2 * NB * 32;    // doubles per group (1024)
    static constexpr int GEN_ITERS = GROUP_SLOTS / THREADS; // covariance elements per thread and group (2)
    static constexpr int RG = RBC / GK;                     // groups of the resident panel (16 -> 128 KB)
    static constexpr int RQ = RBW >= 4 ? 4 : RBW;           // row blocks per ring quad (one cp.async commit group)
    static constexpr int QPC = RBW / RQ;                    // quads per column block
    static constexpr int KA = QD / QPC;                     // the ring runs KA column blocks ahead
    static constexpr int RING_DOUBLES_PER_WARP = QD * RQ * 64;
    static constexpr int NBUF_MAX = 4;                      // streaming group buffers (right of the panel)
    static_assert(NW % NB == 0 && GROUP_SLOTS % THREADS == 0, "generation loop must tile");
    static_assert(QD % QPC == 0 && KA >= 1, "ring depth must be whole column blocks");
    static_assert(RING_DOUBLES_PER_WARP >= T * (1 + RMAX), "the per-warp reduction slice lives in the drained ring");
    static size_t smem_bytes(int D, int nbuf) {
        return ((size_t)NW * RING_DOUBLES_PER_WARP + (size_t)(RG + nbuf) * GROUP_SLOTS + (size_t)D * T + T + T * (1 + RMAX) + 2 +
                34) * sizeof(double) + (size_t)(RG + 2 * NBUF_MAX) * 8;
    }
};

__device__ __forceinline__ void dmma884v(double& c0, double& c1, double a, double b) {      // volatile: issue order = source order
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds128(unsigned addr) {                                   // addr: shared-window address
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
template <int ABL>
__device__ __forceinline__ void cp5(unsigned smem_dst, const void* gmem_src) { if (!(ABL & 2)) cp_async16(smem_dst, gmem_src); }
__device__ __forceinline__ void mbar_wait_par(unsigned bar_s, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT5_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT5_%=;\n\t}" ::"r"(bar_s), "r"(parity) : "memory");
}

// ABL (profiling builds only, GPFO_BUILD_EXPERIMENTS): bit 0 = K(X*,X) elements are constants (no generation arithmetic),
// bit 1 = no L^-1 copies (the ring is never filled), bit 2 = no mbarrier hand-off of the groups (wrong results; timing only)
template <int RBW, int NB, int QD, int KIND, int ABL = 0>
__global__ void __launch_bounds__(512, 1)
contract5_kernel(const GpDev* __restrict__ gpp, const ProgramDev* __restrict__ prog, EvalIO io, int nbuf) {
    using C = Contract5Cfg<RBW, NB, QD>;
    constexpr int T = C::T, NW = C::NW, GK = C::GK, RQ = C::RQ, QPC = C::QPC, RG = C::RG, KA = C::KA;
    constexpr int GS = C::GROUP_SLOTS;
    extern __shared__ __align__(16) double smem[];
    const int D = gpp->D;
    double* ring = smem;                                             // [NW][QD][RQ][64]  private L^-1 rings
    double* kres = ring + (size_t)NW * C::RING_DOUBLES_PER_WARP;     // [RG][GS]          resident panel of K(X*,X)
    double* kstr = kres + (size_t)RG * GS;                           // [nbuf][GS]        streaming groups right of the panel
    double* xc = kstr + (size_t)nbuf * GS;                           // [D][T]
    double* x2 = xc + (size_t)D * T;                                 // [T]
    double* tot = x2 + T;                                            // [T][1+RMAX] running colsum(A^2), A^T V
    double* sbest = tot + T * (1 + RMAX);                            // [2] best value, best index (as bits)
    double* tab32 = sbest + 2;                                       // [32] 2^(i/32) (+2 pad)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tab32 + 34);   // full_res[RG], full_str[4], empty_str[4]
    static_assert(T <= 32, "one epilogue warp");

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const GpDev& gp = *gpp;
    const int M = gp.M, R = gp.R;
    const double variance = gp.variance;
    const double* __restrict__ Xs = gp.Xs;
    const double* __restrict__ xn = gp.xn;
    const double* __restrict__ Vv = gp.V;
    const PackGeom geom = io.geom_o;                  // stream layout of this launch: chunks of RBC row blocks
    const double* __restrict__ packed = io.packed_o;
    double* myring = ring + (size_t)w * C::RING_DOUBLES_PER_WARP + 2 * lane;
    const unsigned myring_s = (unsigned)__cvta_generic_to_shared(myring);
    const unsigned bar_res = (unsigned)__cvta_generic_to_shared(bars);
    const unsigned bar_full = bar_res + 8 * RG, bar_empty = bar_full + 8 * C::NBUF_MAX;

    const int gen_nb = w % NB;
    const int gen_n = 8 * gen_nb + (lane >> 2);
    const int gen_jl = lane & 3;
    // warps w, w+4, w+8, w+12 share a scheduler: two generate before their DMMA work of a step, two after it, so the FP64
    // generation of one pair overlaps the tensor work of the other
    const bool gen_first = ((w >> 2) & 1) == 0;
    const int la = nbuf >= 3 ? 2 : 1;                 // groups are generated `la` steps ahead of use

    if (tid == 0) { sbest[0] = 0.0; reinterpret_cast<long long*>(sbest)[1] = -1; }
    if (tid < 32) tab32[tid] = exp2(tid * (1.0 / 32.0));
    if (tid == 0) {
        for (int b = 0; b < RG; ++b) mbar_init(bar_res + 8 * b, NW);
        for (int b = 0; b < C::NBUF_MAX; ++b) { mbar_init(bar_full + 8 * b, NW); mbar_init(bar_empty + 8 * b, NW); }
    }
    __syncthreads();
    // every streaming buffer starts empty: complete phase 0 of the empty barriers
    if (lane == 0)
        for (int b = 0; b < C::NBUF_MAX; ++b) mbar_arrive(bar_empty + 8 * b);

    unsigned res_par = 0;                             // parity of the resident panel's barriers: one completion per tile
    unsigned full_ph = 0, empty_ph = 0;               // bit b: parity of the next completion to wait for on buffer b
    int gen_buf = 0, use_buf = 0;                     // streaming ring cursors (every thread walks the same sequence)

    const int64_t ntiles = (io.N + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * T;
        __syncthreads();                              // previous tile done: xc, tot and the resident panel may be overwritten
        if (tid < T * (1 + RMAX)) tot[tid] = 0.0;
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

        for (int p = 0; p < geom.nchunks; ++p) {
            const int nr = pg_chunk_rows(geom, p);
            const int row0 = p * geom.rbc;            // first row block of this chunk = first diagonal column block
            const int kp_end = row0 + nr;
            const bool full_chunk = nr == C::RBC;

            double acc[RBW][NB][2];
#pragma unroll
            for (int r = 0; r < RBW; ++r)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) { acc[r][nb][0] = 0.0; acc[r][nb][1] = 0.0; }

            // ---- L^-1 ring.  Block (kp, row block rbl) lives at P(kp) + rbl * 64 doubles, where P advances by nr blocks per dense
            // column block and by nr - kd - 1 per diagonal one (pg_col_base in closed form).  src_next points at column block
            // kp_next = (column block being consumed) + KA, already offset to this warp and lane.
            const double* __restrict__ src_next = packed + pg_chunk_base(geom, p) * 64 + (int64_t)w * 64 + 2 * lane;
            int kp_next = 0;
            int ring_off = 0;                         // doubles: quad slot being consumed (and then refilled)
            // row blocks r < r_hi of this warp exist in this chunk (r * NW + w < nr)
            const int r_hi = nr > w ? (nr - w + NW - 1) / NW : 0;
            auto issue_quad = [&](int qd) {           // ring prologue: quad qd of column block kp_next into slot ring_off
                if (kp_next < kp_end) {
                    const int t = kp_next - row0 - w + (NW - 1);
                    const int n_lo = t > 0 ? t / NW : 0;
                    const unsigned dst = myring_s + ring_off * 8;
#pragma unroll
                    for (int i = 0; i < RQ; ++i)
                        if (qd * RQ + i >= n_lo && qd * RQ + i < r_hi) cp5<ABL>(dst + i * 512, src_next + (qd * RQ + i) * (NW * 64));
                }
                cp_async_commit();
            };
            auto advance_next = [&]() {
                const int kd = kp_next - row0;
                src_next += (kd < 0 ? nr : nr - kd - 1) * 64;
                ++kp_next;
            };
            auto ring_step = [&]() { ring_off = (ring_off == (QD - 1) * RQ * 64) ? 0 : ring_off + RQ * 64; };

            // ---- K(X*,X) groups
            auto gen_elems = [&](int g, double* __restrict__ dst) {
                double dot[C::GEN_ITERS];
                int jj[C::GEN_ITERS];
#pragma unroll
                for (int q = 0; q < C::GEN_ITERS; ++q) {
                    const int ksl = (w + NW * q) / NB;
                    const int j = (g * GK) * 8 + 4 * ksl + gen_jl;
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
                    const int j = (g * GK) * 8 + 4 * ksl + gen_jl;
                    double val;
                    if (ABL & 1) val = 1e-3 * j;
                    else {
                        const double r2 = (-2.0 * dot[q] + __ldg(xn + jj[q])) + x2[gen_n];    // square_dist, expanded form
                        val = kx_branchless<KIND>(variance, r2, tab32);
                    }
                    dst[(ksl * NB + gen_nb) * 32 + lane] = j < M ? val : 0.0;
                }
            };
            auto gen_group = [&](int g) {
                if (g < RG) {                             // resident panel (chunk 0 only): slot g is free since the tile barrier
                    gen_elems(g, kres + (size_t)g * GS);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_res + 8 * g);
                } else {                                  // streaming ring: wait until every warp has released the buffer
                    if (!(ABL & 4)) mbar_wait_par(bar_empty + 8 * gen_buf, (empty_ph >> gen_buf) & 1u);
                    empty_ph ^= 1u << gen_buf;
                    gen_elems(g, kstr + (size_t)gen_buf * GS);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * gen_buf);
                    gen_buf = (gen_buf + 1 == nbuf) ? 0 : gen_buf + 1;
                }
            };

            // One column block: B fragments from the group, A fragments from the ring, up to RBW * 2 * NB DMMAs.
            // Row block r of this warp (global row block r * NW + w of the chunk) is active in column block kp iff
            // r_lo(kp) <= r < r_hi, r_lo(kp) = ceil((kp - row0 - w) / NW) clamped at 0: warp-uniform, zero left of the diagonal,
            // and constant over 16 consecutive diagonal column blocks.
            //
            // mma_static<RLO, NEXT_DENSE>: FULL chunks (r_hi == RBW).  Rows RLO .. RBW-1 are active, rows below RLO do not
            // exist in the instruction stream.  (A predicated-off DMMA still occupies the FP64 tensor pipe for its 16 cycles
            // -- measured: 1.5 x the useful pipe time with `@P DMMA` in the triangles, profiles/r02_v5_first_*.csv -- so the
            // triangles are specialised at compile time and dispatched by a warp-uniform switch instead.)
            // NEXT_DENSE: the column block the ring is refilled for is dense as well: unconditional copies.
            auto mma_static = [&](const double* __restrict__ kbk, auto rlo_tag, auto next_tag) {
                constexpr int RLO = decltype(rlo_tag)::value;
                constexpr bool NEXT_DENSE = decltype(next_tag)::value;
                double bf[2][NB];
                if (RLO < RBW) {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) bf[h][nb] = kbk[(h * NB + nb) * 32];
                }
                // first row block the NEXT column block (kp_next) keeps: RLO or RLO + 1; none at all past the chunk's end
                bool has_next = true, keep_first = true;
                if (!NEXT_DENSE) {
                    has_next = kp_next < kp_end;
                    keep_first = (kp_next - row0 - w) <= RLO * NW;                 // ceil((kp_next - row0 - w) / NW) <= RLO
                }
#pragma unroll
                for (int qd = 0; qd < QPC; ++qd) {
                    if ((qd + 1) * RQ <= RLO) {           // no row block of this quad is left: keep the group count in step
                        cp_async_commit();
                        ring_step();
                        continue;
                    }
                    cp_async_wait<QD - 1>();              // this thread's copies of the quad have landed
                    const unsigned slot = myring_s + ring_off * 8;
                    double2 a[RQ];
#pragma unroll
                    for (int i = 0; i < RQ; ++i)
                        if (qd * RQ + i >= RLO) a[i] = lds128(slot + i * 512);
                    // refill the slot just read for column block kp + KA (the loads above precede these asynchronous
                    // writes in program order; their data returns long before a copy can land)
#pragma unroll
                    for (int i = 0; i < RQ; ++i) {
                        const int r = qd * RQ + i;
                        if (r > RLO) { if (NEXT_DENSE || has_next) cp5<ABL>(slot + i * 512, src_next + r * (NW * 64)); }
                        else if (r == RLO) { if (NEXT_DENSE || (has_next && keep_first)) cp5<ABL>(slot + i * 512, src_next + r * (NW * 64)); }
                    }
                    cp_async_commit();
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < RQ; ++i)
                            if (qd * RQ + i >= RLO) {
#pragma unroll
                                for (int nb = 0; nb < NB; ++nb)
                                    dmma884v(acc[qd * RQ + i][nb][0], acc[qd * RQ + i][nb][1], h == 0 ? a[i].x : a[i].y, bf[h][nb]);
                            }
                    ring_step();
                }
                advance_next();
            };
            auto mma_diag = [&](const double* __restrict__ kbk, int kp) {      // full chunk, diagonal column block: dispatch on r_lo
                const int t = kp - row0 - w + (NW - 1);
                const int r_lo = t > 0 ? t / NW : 0;
                switch (r_lo) {
                    case 0: mma_static(kbk, std::integral_constant<int, 0>(), std::false_type()); break;
                    case 1: mma_static(kbk, std::integral_constant<int, (1 < RBW ? 1 : RBW)>(), std::false_type()); break;
                    case 2: mma_static(kbk, std::integral_constant<int, (2 < RBW ? 2 : RBW)>(), std::false_type()); break;
                    case 3: mma_static(kbk, std::integral_constant<int, (3 < RBW ? 3 : RBW)>(), std::false_type()); break;
                    case 4: mma_static(kbk, std::integral_constant<int, (4 < RBW ? 4 : RBW)>(), std::false_type()); break;
                    case 5: mma_static(kbk, std::integral_constant<int, (5 < RBW ? 5 : RBW)>(), std::false_type()); break;
                    case 6: mma_static(kbk, std::integral_constant<int, (6 < RBW ? 6 : RBW)>(), std::false_type()); break;
                    case 7: mma_static(kbk, std::integral_constant<int, (7 < RBW ? 7 : RBW)>(), std::false_type()); break;
                    default: mma_static(kbk, std::integral_constant<int, RBW>(), std::false_type()); break;
                }
            };
            // ragged last chunk (r_hi < RBW for some warps): generic bounds, predicated (slower; at most one chunk per tile)
            auto mma_ragged = [&](const double* __restrict__ kbk, int kp) {
                double bf[2][NB];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) bf[h][nb] = kbk[(h * NB + nb) * 32];
                int r_lo, n_lo;
                { const int t = kp - row0 - w + (NW - 1); r_lo = t > 0 ? t / NW : 0; }
                { const int t = kp_next - row0 - w + (NW - 1); n_lo = kp_next < kp_end ? (t > 0 ? t / NW : 0) : RBW; }
#pragma unroll
                for (int qd = 0; qd < QPC; ++qd) {
                    if (qd * RQ >= r_hi || (qd + 1) * RQ <= r_lo) {       // no row block of this quad exists here (and none next)
                        cp_async_commit();
                        ring_step();
                        continue;
                    }
                    cp_async_wait<QD - 1>();
                    const unsigned slot = myring_s + ring_off * 8;
                    double2 a[RQ];
#pragma unroll
                    for (int i = 0; i < RQ; ++i) {
                        if (qd * RQ + i >= r_lo && qd * RQ + i < r_hi) a[i] = lds128(slot + i * 512);
                        else a[i] = make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int i = 0; i < RQ; ++i)
                        if (qd * RQ + i >= n_lo && qd * RQ + i < r_hi)
                            cp5<ABL>(slot + i * 512, src_next + (qd * RQ + i) * (NW * 64));
                    cp_async_commit();
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < RQ; ++i) {
                            if (qd * RQ + i >= r_lo && qd * RQ + i < r_hi) {
#pragma unroll
                                for (int nb = 0; nb < NB; ++nb)
                                    dmma884v(acc[qd * RQ + i][nb][0], acc[qd * RQ + i][nb][1], h == 0 ? a[i].x : a[i].y, bf[h][nb]);
                            }
                        }
                    ring_step();
                }
                advance_next();
            };

            const int ngroups = (kp_end + GK - 1) / GK;
            const int g_first = p == 0 ? 0 : RG;      // groups left of it are resident since chunk 0 of this tile

            // ring prologue: KA column blocks ahead
#pragma unroll
            for (int a = 0; a < KA; ++a) {
#pragma unroll
                for (int qd = 0; qd < QPC; ++qd) { issue_quad(qd); ring_step(); }
                advance_next();
            }
            // generation prologue (chunk 0: the first `la` groups)
            for (int g = g_first; g < la && g < ngroups; ++g) gen_group(g);

            for (int g = 0; g < ngroups; ++g) {
                const int gg = g + la;
                const bool do_gen = gg >= g_first && gg < ngroups;
                if (do_gen && gen_first) gen_group(gg);
                // wait for group g
                const double* kb;
                if (g < RG) {
                    if (p == 0 && !(ABL & 4)) mbar_wait_par(bar_res + 8 * g, res_par);
                    kb = kres + (size_t)g * GS + lane;
                } else {
                    if (!(ABL & 4)) mbar_wait_par(bar_full + 8 * use_buf, (full_ph >> use_buf) & 1u);
                    full_ph ^= 1u << use_buf;
                    kb = kstr + (size_t)use_buf * GS + lane;
                }
                const int k0 = g * GK;
                const int k1 = (k0 + GK < kp_end) ? k0 + GK : kp_end;
                if (full_chunk && k1 + KA <= row0) {  // group and its refills in the dense part of a full chunk: no predicates
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp)
                        mma_static(kb + (kp - k0) * (2 * NB * 32), std::integral_constant<int, 0>(), std::true_type());
                } else if (full_chunk) {
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp) mma_diag(kb + (kp - k0) * (2 * NB * 32), kp);
                } else {
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp) mma_ragged(kb + (kp - k0) * (2 * NB * 32), kp);
                }
                if (g >= RG) {                        // release the streaming buffer (its B fragments are consumed)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_empty + 8 * use_buf);
                    use_buf = (use_buf + 1 == nbuf) ? 0 : use_buf + 1;
                }
                if (do_gen && !gen_first) gen_group(gg);
            }
            cp_async_wait<0>();

            // ---- chunk reduction (fixed order -> deterministic).  Warp w's slice red[w][T][1+RMAX] lives at the start of its
            // own, now drained, ring region; the next chunk's copies are issued only after the barriers below.
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
        }
        res_par ^= 1u;
        if (w == 0) {                                     // epilogue: lane n owns candidate n0 + n
            double tot_mv[RMAX];
            const int ln = lane < T ? lane : 0;
#pragma unroll
            for (int ro = 0; ro < RMAX; ++ro) tot_mv[ro] = tot[ln * (1 + RMAX) + 1 + ro];
            double best_v = sbest[0];
            long long best_i = reinterpret_cast<long long*>(sbest)[1];
            gpfo_tile_epilogue<T>(gp, prog, io, n0, lane, tot[ln * (1 + RMAX)], tot_mv, best_v, best_i);
            if (lane == 0) { sbest[0] = best_v; reinterpret_cast<long long*>(sbest)[1] = best_i; }
        }
    }
    __syncthreads();
    if (tid == 0 && io.run_program) {
        io.part_val[blockIdx.x] = sbest[0];
        io.part_idx[blockIdx.x] = reinterpret_cast<long long*>(sbest)[1];
    }
}

template <int RBW, int NB, int QD, int KIND, int ABL = 0>
static inline cudaError_t launch5_kind(cudaStream_t st, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int grid, int D) {
    using C = Contract5Cfg<RBW, NB, QD>;
    int nbuf = C::NBUF_MAX;
    while (nbuf >= 2 && C::smem_bytes(D, nbuf) > (size_t)227 * 1024) --nbuf;
    if (nbuf < 2) return cudaErrorInvalidConfiguration;
    const size_t smem = C::smem_bytes(D, nbuf);
    auto kern = contract5_kernel<RBW, NB, QD, KIND, ABL>;
    static size_t smem_set[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || smem_set[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 16) smem_set[dev] = smem;
    }
    kern<<<grid, C::THREADS, smem, st>>>(d_gp, d_prog, io, nbuf);
    return cudaGetLastError();
}

template <int RBW, int NB, int QD, int ABL = 0>
static inline cudaError_t launch5(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int grid, int D) {
    switch (kind) {
        case GPFO_KERN_RBF: return launch5_kind<RBW, NB, QD, GPFO_KERN_RBF, ABL>(st, d_gp, d_prog, io, grid, D);
        case GPFO_KERN_MATERN52: return launch5_kind<RBW, NB, QD, GPFO_KERN_MATERN52, ABL>(st, d_gp, d_prog, io, grid, D);
        default: return launch5_kind<RBW, NB, QD, GPFO_KERN_MATERN32, ABL>(st, d_gp, d_prog, io, grid, D);
    }
}
