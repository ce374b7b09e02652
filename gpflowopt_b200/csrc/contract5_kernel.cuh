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
//     per-warp work differs by at most one block-step in eight (v2: one in two).  The triangles are specialised at compile
//     time by the first active row block of the warp and dispatched once per run of column blocks with the same first block
//     (mma_diag_run), so their loop bodies are straight-line code like the dense part.
//   * No CTA barrier in the column loop.  Column groups (8 KB each) are handed over through mbarriers: every warp arrives on
//     full[g] after writing its share of group g (generated two groups ahead of use) and waits on it right before the first
//     B-fragment load; the streaming groups right of the panel cycle through a small ring with empty[] barriers.  Warps of a
//     CTA may drift apart by a whole group instead of meeting at a __syncthreads every 8 column blocks.
//   * Tile pipeline: a tile ends with the two barriers of its last chunk reduction and nothing else.  Between them idle
//     threads stage the NEXT tile's (transformed, scaled) candidates into the other half of a double buffer; the epilogue of
//     a tile (moments, acquisition program with its erf / exp, argmax) is run by warp 0 one tile LATER, after it has
//     contributed its share of the first two covariance groups, while the other fifteen warps are already contracting.
//   * K(X*,X) generation: the D-term dot products Xs . xc of a group are 8 x 8 tiles of DMMAs themselves (one tile per warp
//     and group, ceil(D / 4) DMMAs instead of D DFMAs per element) -- on the shared FP64 pipe every generation instruction
//     waits one arbitration round behind the other warps' 16-cycle DMMAs, so the instruction COUNT of the generation is what
//     costs (phase counters: profiles/r02_phase5_*.log); the kernel variance is folded into the 256-entry 2^(i/256) table,
//     the kernel's constants into the squared distance (kx5), and resident groups are generated two at a time (four
//     independent chains per thread).
//   * L^-1 blocks travel global -> shared with cp.async into a per-warp private ring of QD quads of row blocks (each lane
//     copies and re-reads its own 16 bytes: no barrier); the slot just read is refilled for the column block one ahead.
//     One 512-byte block feeds 2 NB DMMAs, so the stream is 4 x (T = 16) the L2 traffic of v2's T = 64 -- 7 TB/s at
//     M = 2048, which the L2 (about 12 TB/s) carries; that is the price of keeping K(X*,X) on chip.
#pragma once
#include <type_traits>
#include "contract_common.cuh"

template <int RBW, int NB, int QD, int NWARPS = 16>
struct Contract5Cfg {
    static constexpr int NW = NWARPS;                       // warps per CTA: 16 (one CTA per SM) or 8 (two CTAs per SM, half-height chunks)
    static constexpr int THREADS = 32 * NW;
    static constexpr int T = 8 * NB;                        // candidates per tile
    static constexpr int RBC = NW * RBW;                    // row blocks per chunk
    static constexpr int GK = NW / NB;                      // column blocks per K(X*,X) group: 8 KB (16 warps) / 4 KB (8 warps)
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
    static_assert(GK * NB == NW, "one 8 x 8 generation tile per warp and group");
    static int dpad(int D) { return (D + 3) / 4 * 4; }     // input dimension rounded up to whole DMMA k-steps (zero padded)
    static size_t smem_bytes(int D, int nbuf) {
        return ((size_t)NW * RING_DOUBLES_PER_WARP + (size_t)(RG + nbuf) * GROUP_SLOTS + 2 * (size_t)dpad(D) * T +
                2 * T * (1 + RMAX) + 2 + 258) * sizeof(double) + (size_t)(RG + 2 * NBUF_MAX) * 8;
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
// K(X*,X) generation arithmetic.  Every FP64 instruction of a generating warp queues behind the other warps' 16-cycle DMMAs on
// the shared pipe, so the forms below are instruction-count-minimal (20 FP64 instructions per Matern52 element, was 24):
//  * the kernel's constants are folded into the squared distance: the caller passes  a = c (r^2 + 1e-12)  (c = 5 / 3 for
//    Matern52 / Matern32, so that sqrt(a) = sqrt(c) r is the exponent itself) or  a = -r^2 / 2  (RBF), built from pre-scaled
//    row norms and tile norms in the same two instructions the plain r^2 took;
//  * exp: table of variance * 2^(i/256) (2 KB), k = round(256 x / ln 2), x = k ln2/256 + r, |r| <= ln2/512, degree-4 Taylor
//    (truncation 3.8e-17 relative); the clamp that keeps 2^(k/256) a normal number is an integer max on the exponent (ALU
//    pipe) instead of an fmax on the argument; the sign of the argument rides on the FMA operands.  9 FP64 instructions.
template <bool NEG>      // NEG: variance * exp(-x) for x >= 0;  else variance * exp(x) for x <= 0
__device__ __forceinline__ double exp_tab256(double x, const double* __restrict__ tabv) {
    const double magic = 6755399441055744.0;                 // 1.5 * 2^52
    const double t = fma(x, NEG ? -369.3299304675746 : 369.3299304675746, magic);      // 256 / ln 2
    const double kf = t - magic;
    double r = fma(kf, -0.00270760617331689, NEG ? -x : x);  // ln2/256, head (31 significant bits: exact product)
    r = fma(kf, -7.453964567463233e-13, r);                  // ln2/256, tail
    double p = 4.1666666666666664e-02;                       // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int lo = __double2loint(t), hi = __double2hiint(t);
    p *= tabv[lo & 255];
    int e = (int)__funnelshift_r((unsigned)lo, (unsigned)hi, 8);   // floor(k / 256): bits 8..39 of the mantissa 2^51 + k
    e = e > -900 ? e : -900;                                 // far tail: ~1e-271 instead of a wrapped exponent
    return __hiloint2double(__double2hiint(p) + (int)((unsigned)e << 20), __double2loint(p));
}
// multipliers of (dot product, row norm, tile norm) and the additive constant that turn them into the argument `a` above
template <int KIND> struct Kx5Scale {
    static constexpr double c = KIND == GPFO_KERN_MATERN52 ? 5.0 : 3.0;
    static constexpr double kd = KIND == GPFO_KERN_RBF ? 1.0 : -2.0 * c;
    static constexpr double kn = KIND == GPFO_KERN_RBF ? -0.5 : c;
    static constexpr double k0 = KIND == GPFO_KERN_RBF ? 0.0 : c * 1e-12;
};
// covariance (variance included through the table) from a = kd * dot + kn * |Xs_j|^2 + (kn * |xc|^2 + k0);
// the forms of contract_common.cuh::kx_branchless with the constants moved into `a`
template <int KIND>
__device__ __forceinline__ double kx5(double a, const double* __restrict__ tabv) {
    if (KIND == GPFO_KERN_RBF) return exp_tab256<false>(fmin(a, 0.0), tabv);
    const double r = fast_sqrt_pos(a);                       // sqrt(c) * euclid_dist
    if (KIND == GPFO_KERN_MATERN52) return fma(1.0 / 3.0, a, r + 1.0) * exp_tab256<true>(r, tabv);     // 1 + sqrt5 r + 5 r^2 / 3
    return (r + 1.0) * exp_tab256<true>(r, tabv);
}
__device__ __forceinline__ void mbar_wait_par(unsigned bar_s, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT5_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT5_%=;\n\t}" ::"r"(bar_s), "r"(parity) : "memory");
}

// ABL (profiling builds only, GPFO_BUILD_EXPERIMENTS): bit 0 = K(X*,X) elements are constants (no generation arithmetic),
// bit 1 = no L^-1 copies (the ring is never filled), bit 2 = no mbarrier hand-off of the groups (wrong results; timing only),
// bit 3 = per-warp phase cycle counters
template <int RBW, int NB, int QD, int KIND, int ABL = 0, int NWARPS = 16>
__global__ void __launch_bounds__(32 * NWARPS, 16 / NWARPS)
contract5_kernel(const GpDev* __restrict__ gpp, const ProgramDev* __restrict__ prog, EvalIO io, int nbuf) {
    using C = Contract5Cfg<RBW, NB, QD, NWARPS>;
    constexpr int T = C::T, NW = C::NW, GK = C::GK, RQ = C::RQ, QPC = C::QPC, RG = C::RG, KA = C::KA;
    constexpr int GS = C::GROUP_SLOTS;
    extern __shared__ __align__(16) double smem[];
    const int D = gpp->D;
    const int DP = (D + 3) / 4 * 4;
    double* ring = smem;                                             // [NW][QD][RQ][64]  private L^-1 rings
    double* kres = ring + (size_t)NW * C::RING_DOUBLES_PER_WARP;     // [RG][GS]          resident panel of K(X*,X)
    double* kstr = kres + (size_t)RG * GS;                           // [nbuf][GS]        streaming groups right of the panel
    double* xcb = kstr + (size_t)nbuf * GS;                          // [2][DP][T]        candidates of this / the next tile
    double* totb = xcb + 2 * (size_t)DP * T;                         // [2][T][1+RMAX]    colsum(A^2), A^T V of this / the previous tile
    double* sbest = totb + 2 * T * (1 + RMAX);                       // [2] best value, best index (as bits)
    double* tabv = sbest + 2;                                        // [256] variance * 2^(i/256) (+2 pad)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tabv + 258);   // full_res[RG], full_str[4], empty_str[4]
    static_assert(T <= 32, "one epilogue warp");

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // Row slot of the warp: its row blocks are r * NW + wr of the chunk.  Warps w, w + 4, w + 8, ... share a scheduler, and in
    // a triangle row block b carries b + 1 column blocks of work, so with wr = w scheduler 3 gets more DMMAs than scheduler 0
    // in every triangle.  Reversing every second group of four makes the slots of a scheduler's warps sum to the same value
    // ({s, 7 - s, 8 + s, 15 - s}).  Measured (profiles/r02_v5_pair_run_lean.log): +0.5 / +0.8 points at M = 2048 / 4096, but
    // -1.5 / -0.4 at M = 512 / 1024, where a tile is a single triangle -- so only tiles of several chunks are dealt this way.
    // (ABL bit 4096, chosen at launch by the chunk count: a run-time choice here costs more than it gains.)
    const int wr = ((ABL & 4096) && ((w >> 2) & 1)) ? (w | 3) - (w & 3) : w;
    const GpDev& gp = *gpp;
    const int M = gp.M, R = gp.R;
    const double* __restrict__ Xs = gp.Xs;
    const double* __restrict__ xn = gp.xn;
    const double* __restrict__ Vv = gp.V;
    const PackGeom geom = io.geom_o;                  // stream layout of this launch: chunks of RBC row blocks
    const double* __restrict__ packed = io.packed_o;
    double* myring = ring + (size_t)w * C::RING_DOUBLES_PER_WARP + 2 * lane;
    const unsigned myring_s = (unsigned)__cvta_generic_to_shared(myring);
    const unsigned bar_res = (unsigned)__cvta_generic_to_shared(bars);
    const unsigned bar_full = bar_res + 8 * RG, bar_empty = bar_full + 8 * C::NBUF_MAX;

    // generation: warp w owns the 8 x 8 tile (column block gen_jb of the group, candidate block gen_nb); after the DMMAs lane l
    // holds the dot products of training point 8 jb + l/4 with candidates 8 nb + 2 (l%4) + {0, 1}
    const int gen_nb = w % NB, gen_jb = w / NB;
    const int gen_n0 = 8 * gen_nb + 2 * (lane & 3);
    // B-fragment slot of element (j, n) inside its group: ((j / 4) * NB + n / 8) * 32 + (n % 8) * 4 + j % 4
    const int gen_slot = ((2 * gen_jb + (lane >> 4)) * NB + gen_nb) * 32 + 8 * (lane & 3) + ((lane >> 2) & 3);
    // warps w, w+4, w+8, w+12 share a scheduler: two generate before their DMMA work of a step, two after it
    const bool gen_first = (ABL & 16) ? true : ((ABL & 64) ? false : ((w >> 2) & 1) == 0);
    const int la = (ABL & 32) ? 1 : (nbuf >= 3 ? 2 : 1);   // groups are generated `la` steps ahead of use

    if (tid == 0) { sbest[0] = 0.0; reinterpret_cast<long long*>(sbest)[1] = -1; }
    for (int e = tid; e < 256; e += C::THREADS) tabv[e] = gp.variance * exp2(e * (1.0 / 256.0));
    if (tid < 2 * T * (1 + RMAX)) totb[tid] = 0.0;
    if (tid == 0) {
        for (int b = 0; b < RG; ++b) mbar_init(bar_res + 8 * b, NW);
        for (int b = 0; b < C::NBUF_MAX; ++b) { mbar_init(bar_full + 8 * b, NW); mbar_init(bar_empty + 8 * b, NW); }
    }
    // candidates of tile `tile` -> xcb[buf]: input transform (transforms.py:102-103), division by the lengthscales, zero padding
    // of the input dimension up to DP.  One element per thread and pass; the caller provides the barrier that publishes them.
    auto stage_candidates = [&](int64_t tile, int buf, int first_thread, int nthreads) {
        const int64_t n0 = tile * T;
        double* xc = xcb + (size_t)buf * DP * T;
        for (int e = tid - first_thread; e >= 0 && e < T * DP; e += nthreads) {
            const int n = e / DP, d = e - n * DP;
            double v = 0.0;
            if (d < D) {
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
                v = xt / gp.ell[d];
            }
            xc[d * T + n] = v;
        }
    };
    const int64_t ntiles = (io.N + T - 1) / T;
    if ((int64_t)blockIdx.x < ntiles) stage_candidates(blockIdx.x, 0, 0, C::THREADS);
    __syncthreads();
    // every streaming buffer starts empty: complete phase 0 of the empty barriers
    if (lane == 0)
        for (int b = 0; b < C::NBUF_MAX; ++b) mbar_arrive(bar_empty + 8 * b);

    constexpr bool TIM = (ABL & 8) != 0;              // profiling build: per-warp cycle counters per phase
    long long t_all = 0, t_gen = 0, t_mma = 0, t_wait = 0, t_pro = 0, t_red = 0, t_epi = 0, t_mark = 0;
    if (TIM) t_all = clock64();
    unsigned res_par = 0;                             // parity of the resident panel's barriers: one completion per tile
    unsigned full_ph = 0, empty_ph = 0;               // bit b: parity of the next completion to wait for on buffer b
    int gen_buf = 0, use_buf = 0;                     // streaming ring cursors (every thread walks the same sequence)

    // tile epilogue, run by warp 0 one tile late: lane n owns candidate n0 + n of the tile whose sums sit in totb[buf]
    auto run_epilogue = [&](int64_t n0, int buf) {
        const double* tot = totb + buf * T * (1 + RMAX);
        double tot_mv[RMAX];
        const int ln = lane < T ? lane : 0;
#pragma unroll
        for (int ro = 0; ro < RMAX; ++ro) tot_mv[ro] = tot[ln * (1 + RMAX) + 1 + ro];
        double best_v = sbest[0];
        long long best_i = reinterpret_cast<long long*>(sbest)[1];
        gpfo_tile_epilogue<T>(gp, prog, io, n0, lane, tot[ln * (1 + RMAX)], tot_mv, best_v, best_i);
        if (lane == 0) { sbest[0] = best_v; reinterpret_cast<long long*>(sbest)[1] = best_i; }
    };

    int par = 0;                                      // which half of xcb / totb belongs to the current tile
    int64_t prev_n0 = -1;                             // first candidate of the tile whose epilogue is still due
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * T;
        const double* xc = xcb + (size_t)par * DP * T;
        double* tot = totb + par * T * (1 + RMAX);
        if (TIM) t_mark = clock64();
        // squared norms of this thread's two candidates (its generation tile's columns)
        double x2a = 0.0, x2b = 0.0;
        for (int d = 0; d < D; ++d) {
            const double2 v = *reinterpret_cast<const double2*>(xc + d * T + gen_n0);
            x2a = fma(v.x, v.x, x2a);
            x2b = fma(v.y, v.y, x2b);
        }
        x2a = fma(Kx5Scale<KIND>::kn, x2a, Kx5Scale<KIND>::k0);      // the tile's share of the covariance argument (kx5)
        x2b = fma(Kx5Scale<KIND>::kn, x2b, Kx5Scale<KIND>::k0);
        if (TIM) t_pro += clock64() - t_mark;

        for (int p = 0; p < geom.nchunks; ++p) {
            const int nr = pg_chunk_rows(geom, p);
            const int row0 = p * geom.rbc;            // first row block of this chunk = first diagonal column block
            const int kp_end = row0 + nr;
            // A last chunk with fewer than RBC row blocks runs the same compile-time specialised paths as a full one: the row blocks
            // a warp does not have (r >= r_hi) are ZERO rows -- their ring slots are cleared below and never refilled -- so their
            // DMMAs add nothing.  (The generic predicated path this replaces ran at 63-65 % of peak at M = 2000 / 2040, where the
            // ragged chunk is three quarters of the work; profiles/r02_v5_ragged_chunks.log.)
            // (ABL bit 16384 = SHORT, chosen at launch when the row blocks do not fill the last chunk: the guards this needs in the
            // refill code cost 4-5 points when compiled into the kernel for whole chunks, which therefore stays as it was.)
            constexpr bool SHORT = (ABL & 16384) != 0;
            const bool full_chunk = SHORT ? true : (nr == C::RBC);
            const bool last_chunk = p + 1 == geom.nchunks;

            double acc[RBW][NB][2];                  // zeroed right before the column loop (after the deferred epilogue)
#include "contract5_mma.inc"                          // L^-1 ring + column-block bodies (shared with the backward kernel)
            // ---- K(X*,X) groups: this warp's 8 x 8 tile of dot products on the tensor pipe, then the covariance of its two elements
            auto gen_elems = [&](int g, double* __restrict__ dst) {
                const int jrow = (g * GK + gen_jb) * 8 + (lane >> 2);       // training point of this lane's A-fragment row = of its outputs
                const int jc = jrow < M ? jrow : M - 1;                     // clamp: keeps the loads in range, result masked below
                double d0 = 0.0, d1 = 0.0;
                if (!(ABL & 1)) {
                    const double* __restrict__ xrow = Xs + (size_t)jc * D + (lane & 3);
                    const double* __restrict__ xcol = xc + (lane & 3) * T + 8 * gen_nb + (lane >> 2);
                    for (int k = 0; k < DP; k += 4) {
                        const double a = (k + (lane & 3) < D) ? __ldg(xrow + k) : 0.0;
                        dmma884(d0, d1, a, xcol[k * T]);
                    }
                }
                double v0, v1;
                if (ABL & 1) { v0 = 1e-3 * jrow; v1 = v0; }
                else {
                    const double xnj = Kx5Scale<KIND>::kn * __ldg(xn + jc);
                    v0 = kx5<KIND>(fma(Kx5Scale<KIND>::kd, d0, xnj) + x2a, tabv);          // square_dist, expanded form, scaled
                    v1 = kx5<KIND>(fma(Kx5Scale<KIND>::kd, d1, xnj) + x2b, tabv);
                }
                const bool in = jrow < M;
                dst[gen_slot] = in ? v0 : 0.0;
                dst[gen_slot + 4] = in ? v1 : 0.0;
            };
            // two resident groups at once: four independent covariance chains per thread instead of two (the chain of ~25
            // dependent FP64 instructions per element is what a generating warp waits on: `wait` is the top stall reason)
            auto gen_pair = [&](int g) {
                const int jrow0 = (g * GK + gen_jb) * 8 + (lane >> 2), jrow1 = jrow0 + 8 * GK;
                const int jc0 = jrow0 < M ? jrow0 : M - 1, jc1 = jrow1 < M ? jrow1 : M - 1;
                double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
                const double* __restrict__ xrow0 = Xs + (size_t)jc0 * D + (lane & 3);
                const double* __restrict__ xrow1 = Xs + (size_t)jc1 * D + (lane & 3);
                const double* __restrict__ xcol = xc + (lane & 3) * T + 8 * gen_nb + (lane >> 2);
                for (int k = 0; k < DP; k += 4) {
                    const bool ink = k + (lane & 3) < D;
                    const double a0 = ink ? __ldg(xrow0 + k) : 0.0, a1 = ink ? __ldg(xrow1 + k) : 0.0;
                    const double b = xcol[k * T];
                    dmma884(d0, d1, a0, b);
                    dmma884(e0, e1, a1, b);
                }
                constexpr double KD = Kx5Scale<KIND>::kd;
                const double xn0 = Kx5Scale<KIND>::kn * __ldg(xn + jc0), xn1 = Kx5Scale<KIND>::kn * __ldg(xn + jc1);
                const double v0 = kx5<KIND>(fma(KD, d0, xn0) + x2a, tabv), v1 = kx5<KIND>(fma(KD, d1, xn0) + x2b, tabv);
                const double u0 = kx5<KIND>(fma(KD, e0, xn1) + x2a, tabv), u1 = kx5<KIND>(fma(KD, e1, xn1) + x2b, tabv);
                double* __restrict__ dst0 = kres + (size_t)g * GS;
                dst0[gen_slot] = jrow0 < M ? v0 : 0.0;
                dst0[gen_slot + 4] = jrow0 < M ? v1 : 0.0;
                dst0[GS + gen_slot] = jrow1 < M ? u0 : 0.0;
                dst0[GS + gen_slot + 4] = jrow1 < M ? u1 : 0.0;
                __syncwarp();
                if (lane == 0) { mbar_arrive(bar_res + 8 * g); mbar_arrive(bar_res + 8 * (g + 1)); }
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

            const int ngroups = (kp_end + GK - 1) / GK;
            const int g_first = p == 0 ? 0 : RG;      // groups left of it are resident since chunk 0 of this tile

            if (SHORT && nr != C::RBC) {              // clear the ring slots of this warp's missing row blocks (all stages)
#pragma unroll
                for (int sq = 0; sq < QD; ++sq)
#pragma unroll
                    for (int i = 0; i < RQ; ++i)
                        if ((sq % QPC) * RQ + i >= r_hi) *reinterpret_cast<double2*>(myring + (sq * RQ + i) * 64) = make_double2(0.0, 0.0);
            }
            // ring prologue: KA column blocks ahead
#pragma unroll
            for (int a = 0; a < KA; ++a) {
#pragma unroll
                for (int qd = 0; qd < QPC; ++qd) { issue_quad(qd); ring_step(); }
                advance_next();
            }
            // generation prologue (chunk 0: the first `la` groups)
            if (TIM) t_mark = clock64();
            constexpr bool PAIR = (ABL & 256) == 0;       // (bit 256, profiling builds: one group at a time, as before)
            // PAIR: resident groups are generated two at a time, at even steps (needs the default lookahead of two groups and
            // an even group count, i.e. whole pairs inside the panel; otherwise the single-group schedule below)
            const bool pair = PAIR && p == 0 && la == 2 && (ngroups & 1) == 0;
            // WIDE steps (instantiations for single-chunk tiles): NB groups = NW column blocks per step -- one run of constant
            // first row block: one barrier round, one dispatch, NB / 2 generated pairs per step.  Measured
            // (profiles/r02_v5_pair_run_lean.log): M = 512 (T = 32: four groups per step) 58.3 -> 62.9 %, M = 1024 (T = 16:
            // two groups per step) 72.0 -> 72.5 %; 32 column blocks per step at M = 1024: 71.9 %; in the instantiation for
            // tiles of several chunks (bit 4096) wide steps cost 1.6 points and are off.
            constexpr int GSTEP = NB >= 2 ? NB : 2;
            const bool wide = NB >= 2 && pair && !(ABL & (2048 | 4096)) && (ngroups % GSTEP) == 0;
            static_assert(GSTEP == 2 || GSTEP == 4, "wide steps: one or two generated pairs per step");
            if (wide) { gen_pair(0); if (GSTEP == 4) gen_pair(2); }
            else if (pair) gen_pair(0);
            else for (int g = g_first; g < la && g < ngroups; ++g) gen_group(g);
            if (TIM) { const long long t = clock64(); t_gen += t - t_mark; t_mark = t; }
            // the previous tile's epilogue: warp 0 has delivered its share of the first groups, the other warps are contracting
            if (p == 0 && w == 0 && prev_n0 >= 0) run_epilogue(prev_n0, par ^ 1);
            if (TIM) t_epi += clock64() - t_mark;

#pragma unroll
            for (int r = 0; r < RBW; ++r)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) { acc[r][nb][0] = 0.0; acc[r][nb][1] = 0.0; }

            const int gstep = wide ? GSTEP : 1;
            for (int g = 0; g < ngroups; g += gstep) {
                const int gg = wide ? g + GSTEP : g + la;
                const bool do_gen = pair ? false : (gg >= g_first && gg < ngroups);
                const bool do_pair = pair && (g & 1) == 0 && gg < ngroups;
                if (TIM) t_mark = clock64();
                if (do_pair && gen_first) {
                    gen_pair(gg);
                    if (GSTEP == 4 && wide) gen_pair(gg + 2);
                }
                if (do_gen && gen_first) gen_group(gg);
                if (TIM) { const long long t = clock64(); t_gen += t - t_mark; t_mark = t; }
                // wait for group g
                const double* kb;
                if (g < RG) {
                    if (p == 0 && !(ABL & 4)) {
                        mbar_wait_par(bar_res + 8 * g, res_par);
                        if (wide) {
#pragma unroll
                            for (int q = 1; q < GSTEP; ++q) mbar_wait_par(bar_res + 8 * (g + q), res_par);
                        }
                    }
                    kb = kres + (size_t)g * GS + lane;
                } else {
                    if (!(ABL & 4)) mbar_wait_par(bar_full + 8 * use_buf, (full_ph >> use_buf) & 1u);
                    full_ph ^= 1u << use_buf;
                    kb = kstr + (size_t)use_buf * GS + lane;
                }
                if (TIM) { const double probe = kb[0]; if (probe == 123.456) t_gen += 1; const long long t = clock64(); t_wait += t - t_mark; t_mark = t; }
                const int k0 = g * GK;
                const int k1 = (k0 + gstep * GK < kp_end) ? k0 + gstep * GK : kp_end;
                if (full_chunk && k1 + KA <= row0) {  // group and its refills in the dense part of a full chunk: no predicates
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp)
                        mma_static(kb + (kp - k0) * (2 * NB * 32), std::integral_constant<int, 0>(), std::true_type());
                } else if (full_chunk) {
                    if (ABL & 1024) {                 // (profiling build: the per-column-block dispatch this replaced)
#pragma unroll 1
                        for (int kp = k0; kp < k1; ++kp) mma_diag(kb + (kp - k0) * (2 * NB * 32), kp);
                    } else mma_diag_run(kb, k0, k1);
                } else {
#pragma unroll 1
                    for (int kp = k0; kp < k1; ++kp) mma_ragged(kb + (kp - k0) * (2 * NB * 32), kp);
                }
                if (g >= RG) {                        // release the streaming buffer (its B fragments are consumed)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_empty + 8 * use_buf);
                    use_buf = (use_buf + 1 == nbuf) ? 0 : use_buf + 1;
                }
                if (TIM) { const long long t = clock64(); t_mma += t - t_mark; t_mark = t; }
                if (do_gen && !gen_first) gen_group(gg);
                if (do_pair && !gen_first) {
                    gen_pair(gg);
                    if (GSTEP == 4 && wide) gen_pair(gg + 2);
                }
                if (TIM) { const long long t = clock64(); t_gen += t - t_mark; }
            }
            cp_async_wait<0>();
            if (TIM) t_mark = clock64();
            if (ABL & 128) {
                // stored-A gradient pass: park A = L^-1 Kx of this chunk for the backward launch, which walks 64-candidate
                // tiles: reversed row order, B-fragment group layout of contract2_kernel (8 column blocks of 8 candidate
                // blocks per group); this tile fills candidate blocks NB * (tile % (8 / NB)) .. of 64-candidate tile tile / (8 / NB)
                // (bit 512: the tall-chunk backward kernel reads 16-candidate tiles: NB candidate blocks per group, own tile stride)
                constexpr bool NATIVE = (ABL & 512) != 0;
                constexpr int TPB = NATIVE ? 1 : 8 / NB;    // tiles of this kernel per tile of the backward launch
                constexpr int NBB = NATIVE ? NB : 8;        // candidate blocks per group in the scratch layout
                double* const at = io.a_scratch + (size_t)(tile / TPB) * io.a_stride;
                const int nb0 = (int)(tile % TPB) * NB;
                const int i8 = 7 - (lane >> 2);
#pragma unroll
                for (int r = 0; r < RBW; ++r) {
                    const int rbl = r * NW + wr;
                    if (rbl < nr) {
                        const int rbrev = geom.nrb - 1 - (row0 + rbl);
                        double* const o = at + (size_t)(rbrev >> 3) * (KPC * 2 * NBB * 32) +
                                          (((rbrev & 7) * 2 + (i8 >> 2)) * NBB + nb0) * 32 + (i8 & 3) + 8 * (lane & 3);
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) {
                            __stcg(o + nb * 32, acc[r][nb][0]);
                            __stcg(o + nb * 32 + 4, acc[r][nb][1]);
                        }
                    }
                }
            }

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
                        const int rbl = r * NW + wr;
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
            __syncthreads();                              // every warp is past its column loop: xcb[par] and the panel are dead
            if (tid < T * (1 + RMAX)) {                  // thread (n, k): fixed warp order -> deterministic
                const int n = tid / (1 + RMAX), k = tid - n * (1 + RMAX);
                if (k <= R) {
                    double a = p == 0 ? 0.0 : tot[tid];
                    for (int ww = 0; ww < NW; ++ww) a += red[ww * RS + n * (1 + RMAX) + k];
                    tot[tid] = a;
                }
            } else if (last_chunk && tile + gridDim.x < ntiles) {
                // the otherwise idle threads stage the next tile's candidates into the other buffer; the barrier below publishes them
                stage_candidates(tile + gridDim.x, par ^ 1, T * (1 + RMAX), C::THREADS - T * (1 + RMAX));
            }
            __syncthreads();
            if (TIM) { if (*reinterpret_cast<volatile double*>(tabv) == 123.456) t_gen += 1; t_red += clock64() - t_mark; }
        }
        res_par ^= 1u;
        prev_n0 = n0;
        par ^= 1;
    }
    if (TIM) t_mark = clock64();
    if (w == 0 && prev_n0 >= 0) run_epilogue(prev_n0, par ^ 1);      // the last tile's epilogue
    if (TIM) t_epi += clock64() - t_mark;
    if (TIM && lane == 0 && io.debug) {
        long long* o = io.debug + ((size_t)blockIdx.x * NW + w) * 8;
        o[0] = clock64() - t_all; o[1] = t_gen; o[2] = t_mma; o[3] = t_wait; o[4] = t_pro; o[5] = t_red; o[6] = t_epi;
    }
    __syncthreads();
    if (tid == 0 && io.run_program) {
        io.part_val[blockIdx.x] = sbest[0];
        io.part_idx[blockIdx.x] = reinterpret_cast<long long*>(sbest)[1];
    }
}

template <int RBW, int NB, int QD, int KIND, int ABL = 0, int NWARPS = 16>
static inline cudaError_t launch5_kind(cudaStream_t st, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int grid, int D) {
    using C = Contract5Cfg<RBW, NB, QD, NWARPS>;
    int nbuf = C::NBUF_MAX;
    // (8 warps: two CTAs share the SM's 227 KB, and each CTA's 1 KB reserve)
    const size_t budget = NWARPS == 16 ? (size_t)227 * 1024 : (size_t)(227 * 1024) / 2 - 1024;
    while (nbuf > 2 && C::smem_bytes(D, nbuf) > budget) --nbuf;
    while (nbuf >= 2 && C::smem_bytes(D, nbuf) > (size_t)227 * 1024) --nbuf;
    if (nbuf < 2) return cudaErrorInvalidConfiguration;
    const size_t smem = C::smem_bytes(D, nbuf);
    auto kern = contract5_kernel<RBW, NB, QD, KIND, ABL, NWARPS>;
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

template <int RBW, int NB, int QD, int ABL = 0, int NWARPS = 16>
static inline cudaError_t launch5(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int grid, int D) {
    switch (kind) {
        case GPFO_KERN_RBF: return launch5_kind<RBW, NB, QD, GPFO_KERN_RBF, ABL, NWARPS>(st, d_gp, d_prog, io, grid, D);
        case GPFO_KERN_MATERN52: return launch5_kind<RBW, NB, QD, GPFO_KERN_MATERN52, ABL, NWARPS>(st, d_gp, d_prog, io, grid, D);
        default: return launch5_kind<RBW, NB, QD, GPFO_KERN_MATERN32, ABL, NWARPS>(st, d_gp, d_prog, io, grid, D);
    }
}
