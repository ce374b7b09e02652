// Device helpers shared by the contraction kernels (contract.cu: register-prefetch v1, contract2.cu: cp.async v2).
#pragma once
#include "gpfo_internal.cuh"

#define KPC 8            // column blocks (of 8 training points) per generated Kx group
#define RMAX 4           // max output columns of one GP handled by the kernel

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double2 ldg_stream(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ double kx_from_r2(int kind, double variance, double r2) {
    if (kind == GPFO_KERN_RBF) return variance * exp(-r2 / 2.0);
    const double r = sqrt(r2 + 1e-12);
    if (kind == GPFO_KERN_MATERN52) {
        const double s5 = 2.2360679774997898;
        return variance * (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * exp(-s5 * r);
    }
    const double s3 = 1.7320508075688772;
    return variance * (1.0 + s3 * r) * exp(-s3 * r);
}


// Tile epilogue executed by warp 0: lane n (< T) owns candidate n0 + n and holds colsum(A^2) in tot_ss and
// A^T V in tot_mv[].  Un-scales the moments (transforms.py:112-140 for a diagonal map), optionally stores them,
// runs the acquisition program and folds the tile's best (value, index) into (best_v, best_i).
template <int T>
__device__ __forceinline__ void gpfo_tile_epilogue(const GpDev& gp, const ProgramDev* __restrict__ prog,
                                                   const EvalIO& io, int64_t n0, int lane, double tot_ss,
                                                   const double (&tot_mv)[RMAX], double& best_v, long long& best_i) {
    const int R = gp.R;
    double val = 0.0;
    long long idx = -1;
    const int64_t n = n0 + lane;
    if (lane < T && n < io.N) {
        double mu[GPFO_MAX_SLOTS], va[GPFO_MAX_SLOTS];
        const double fvar = gp.variance - tot_ss;                      // Kdiag - reduce_sum(square(A), 0)
#pragma unroll
        for (int ro = 0; ro < RMAX; ++ro) {
            if (ro < R) {
                const int s = gp.slot0 + ro;
                const double m_o = (tot_mv[ro] - gp.out_shift[ro]) / gp.out_scale[ro];
                const double v_o = fvar / (gp.out_scale[ro] * gp.out_scale[ro]);
                mu[s] = m_o; va[s] = v_o;
                if (io.write_moments) {
                    io.mom_mean[(size_t)s * io.mstride + n] = m_o;
                    io.mom_var[(size_t)s * io.mstride + n] = v_o;
                }
            }
        }
        if (io.run_program) {
            for (int s = 0; s < prog->nslots; ++s) {
                if (s < gp.slot0 || s >= gp.slot0 + R) {
                    mu[s] = io.mom_mean[(size_t)s * io.mstride + n];
                    va[s] = io.mom_var[(size_t)s * io.mstride + n];
                }
            }
            val = gpfo_run_program<false>(*prog, mu, va, 0.0, nullptr, nullptr, nullptr, nullptr);
            if (io.values) io.values[n] = val;
            idx = io.idx0 + n;
        }
    }
    if (io.run_program) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, val, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (gpfo_better(ov, oi, val, idx)) { val = ov; idx = oi; }
        }
        if (gpfo_better(val, idx, best_v, best_i)) { best_v = val; best_i = idx; }
    }
}

// ---- helpers shared by the cp.async kernels (contract2.cu, contract3.cu) ----------------------------------------
template <int NW, int RBW, int NB, int S, int KPG = KPC>
struct Contract2Cfg {
    static constexpr int KP = KPG;                 // column blocks per generated K(X*,X) group
    static constexpr int THREADS = 32 * NW;
    static constexpr int RBC = NW * RBW;
    static constexpr int T = 8 * NB;
    static constexpr int GROUP_SLOTS = KPG * 2 * NB * 32;
    static constexpr int GEN_ITERS = GROUP_SLOTS / THREADS;
    static constexpr int RING_DOUBLES_PER_WARP = S * RBW * 64;
    static_assert(NW % NB == 0, "per-thread candidate must be fixed");
    static_assert(GROUP_SLOTS % THREADS == 0, "gen loop must tile");
};

__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void* gmem_src) {     // smem_dst: shared-window address
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
// CTA-scope split barrier on an mbarrier: arrive (release) as soon as this thread's shared-memory writes are done,
// wait (acquire) only right before the first dependent read.  `phase` is the per-thread parity bit.
__device__ __forceinline__ void mbar_init(unsigned bar_s, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar_s) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar_s) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar_s, unsigned& phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(bar_s), "r"(phase) : "memory");
    phase ^= 1u;
}
__device__ __forceinline__ void cp_async8(unsigned smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The FP64 ALU and the DMMA unit share one issue pipe on B200 (profiles/r01_fp64_pipes.log: 37.0 TF/s DMMA alone,
// 24.4 DFMA alone, 26.6 + 6.7 together), so every FP64 instruction spent on K(X*, X) is taken from the contraction.
// The helpers below are therefore op-count-minimal and branch-free.

// sqrt(x) for normal positive x: MUFU.RSQ64H seed (2^-22), one cubically convergent step on 1/sqrt (-> 2^-66),
// then x * y.  About 1 ulp; no IEEE slow path.  6 FP64 ops.
__device__ __forceinline__ double fast_sqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y * y, 1.0);
    y = fma(y * e, fma(e, 0.375, 0.5), y);
    return x * y;
}

// exp(x) for x <= 0: clamp at -708 (~1e-308 instead of a denormal / 0); k = round(32 x / ln 2), x = k ln2/32 + r
// with |r| <= ln2/64; exp(x) = 2^(k>>5) * 2^((k&31)/32) * P6(r), P6 = degree-6 Taylor (truncation 3.5e-18 relative);
// the 32-entry table of 2^(i/32) lives in shared memory.  12 FP64 ops.
__device__ __forceinline__ double fast_exp_neg(double x, const double* __restrict__ tab32) {
    x = fmax(x, -708.0);
    const double magic = 6755399441055744.0;                 // 1.5 * 2^52
    const double t = fma(x, 46.166241308446828, magic);      // 32 / ln 2
    const double kf = t - magic;
    double r = fma(kf, -2.1660849392498290e-02, x);          // ln2/32 (head)
    r = fma(kf, -7.2470213052434896e-19, r);                 // ln2/32 (tail)
    double p = 1.3888888888888889e-03;                       // 1/6!
    p = fma(p, r, 8.3333333333333332e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int k = __double2loint(t);                         // low word of t holds k (two's complement)
    p *= tab32[k & 31];
    return __hiloint2double(__double2hiint(p) + ((k >> 5) << 20), __double2loint(p));
}

template <int KIND>
__device__ __forceinline__ double kx_branchless(double variance, double r2, const double* __restrict__ tab32) {
    if (KIND == GPFO_KERN_RBF) return variance * fast_exp_neg(-0.5 * fmax(r2, 0.0), tab32);
    const double x = r2 + 1e-12;                             // euclid_dist: sqrt(r2 + 1e-12); square(r) == x to 1 ulp
    const double r = fast_sqrt_pos(x);
    if (KIND == GPFO_KERN_MATERN52) {
        const double s5 = 2.2360679774997898;
        return (variance * fma(5.0 / 3.0, x, fma(s5, r, 1.0))) * fast_exp_neg(-s5 * r, tab32);
    }
    const double s3 = 1.7320508075688772;
    return (variance * fma(s3, r, 1.0)) * fast_exp_neg(-s3 * r, tab32);
}

// d k / d r^2 for the gradient pass: RBF -k/2; Matern52 -(5/6) s2 (1 + sqrt5 r) e^{-sqrt5 r};
// Matern32 -(3/2) s2 e^{-sqrt3 r}, with r = sqrt(r^2 + 1e-12) as in euclid_dist.
template <int KIND>
__device__ __forceinline__ double dk_dr2_branchless(double variance, double r2, const double* __restrict__ tab32) {
    if (KIND == GPFO_KERN_RBF) return (-0.5 * variance) * fast_exp_neg(-0.5 * fmax(r2, 0.0), tab32);
    const double r = fast_sqrt_pos(r2 + 1e-12);
    if (KIND == GPFO_KERN_MATERN52) {
        const double s5 = 2.2360679774997898;
        return ((-5.0 / 6.0) * variance * fma(s5, r, 1.0)) * fast_exp_neg(-s5 * r, tab32);
    }
    const double s3 = 1.7320508075688772;
    return (-1.5 * variance) * fast_exp_neg(-s3 * r, tab32);
}

