// Gradient pass, step 2 of 3: per candidate, evaluate the acquisition program in forward mode over
// the moment slots -> value, d acq / d mean_s, d acq / d var_s, and the block argmax.  Replaces the part of
// tf.gradients(acq, [Xcand]) (gpflowopt/acquisition/acquisition.py:256-257) that differentiates the epilogue
// (ei.py:70-79, poi.py:63-67, pof.py:81-85, lcb.py:54-57, acquisition.py:360-362); the chain through the GP
// prediction is the contraction kernel's gradient mode (contract2.cu).
#include "gpfo_internal.cuh"

#define PG_THREADS 256

__global__ void __launch_bounds__(PG_THREADS) program_grad_kernel(const ProgramDev* __restrict__ prog, EvalIO io) {
    __shared__ double sv[PG_THREADS / 32];
    __shared__ long long si[PG_THREADS / 32];
    double best_v = 0.0;
    long long best_i = -1;
    const int ns = prog->nslots;
    for (int64_t n = (int64_t)blockIdx.x * PG_THREADS + threadIdx.x; n < io.N; n += (int64_t)gridDim.x * PG_THREADS) {
        double mu[GPFO_MAX_SLOTS], va[GPFO_MAX_SLOTS], dmu[GPFO_MAX_SLOTS], dva[GPFO_MAX_SLOTS];
        for (int s = 0; s < ns; ++s) {
            mu[s] = io.mom_mean[(size_t)s * io.mstride + n];
            va[s] = io.mom_var[(size_t)s * io.mstride + n];
        }
        // the VALUE comes from the same instantiation the plain evaluate path uses, so evaluate(X) == evaluate_with_gradients(X)[0]
        // bit for bit (reference testing/unit/test_implementations.py:33); the forward-mode pass supplies the partials only
        const double val = gpfo_run_program<false>(*prog, mu, va, 0.0, nullptr, nullptr, nullptr, nullptr);
        gpfo_run_program<true>(*prog, mu, va, 0.0, nullptr, nullptr, dmu, dva);
        for (int s = 0; s < ns; ++s) {
            io.cmu[(size_t)s * io.mstride + n] = dmu[s];
            io.cva[(size_t)s * io.mstride + n] = dva[s];
        }
        if (io.values) io.values[n] = val;
        if (gpfo_better(val, io.idx0 + n, best_v, best_i)) { best_v = val; best_i = io.idx0 + n; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best_v, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (gpfo_better(ov, oi, best_v, best_i)) { best_v = ov; best_i = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best_v; si[threadIdx.x >> 5] = best_i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < PG_THREADS / 32; ++k)
            if (gpfo_better(sv[k], si[k], best_v, best_i)) { best_v = sv[k]; best_i = si[k]; }
        io.part_val[blockIdx.x] = best_v;
        io.part_idx[blockIdx.x] = best_i;
    }
}

int gpfo_program_grad_grid(int64_t N, int num_sms) {
    const int64_t blocks = (N + PG_THREADS - 1) / PG_THREADS;
    const int64_t cap = (int64_t)num_sms * 4;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

cudaError_t gpfo_program_grad_launch(cudaStream_t st, const ProgramDev* d_prog, EvalIO io, int grid) {
    program_grad_kernel<<<grid, PG_THREADS, 0, st>>>(d_prog, io);
    return cudaGetLastError();
}
