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
                    io.mom_mean[(size_t)s * io.N + n] = m_o;
                    io.mom_var[(size_t)s * io.N + n] = v_o;
                }
            }
        }
        if (io.run_program) {
            for (int s = 0; s < prog->nslots; ++s) {
                if (s < gp.slot0 || s >= gp.slot0 + R) {
                    mu[s] = io.mom_mean[(size_t)s * io.N + n];
                    va[s] = io.mom_var[(size_t)s * io.N + n];
                }
            }
            val = gpfo_run_program<false>(*prog, mu, va, 0.0, nullptr, nullptr, nullptr, nullptr);
            if (io.values) io.values[n] = val;
            idx = n;
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
