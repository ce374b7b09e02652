// Dependent-issue latency of DMMA.8x8x4 (mma.sync.m8n8k4.f64) on sm_100a: throughput of W warps per scheduler, each
// cycling over K independent accumulator pairs.  K small -> the warp waits for its own previous result.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/dmma_latency tools/dmma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int K>
__global__ void chain(int iters, double* out, long long* cyc) {
    double c[K][2];
#pragma unroll
    for (int k = 0; k < K; ++k) { c[k][0] = threadIdx.x; c[k][1] = k; }
    const double a = 1e-3 * threadIdx.x, b = 1.0 + 1e-9 * threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) dmma(c[k][0], c[k][1], a, b);
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int K>
static void run(int warps_per_sched, double* out, long long* cyc) {
    const int iters = 4096 * 32 / K;
    chain<K><<<148, 128 * warps_per_sched>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per = (double)h / ((double)iters * K);
    printf("  K=%2d chains, %d warp(s)/scheduler: %7.1f cycles per DMMA of one warp  -> pipe %5.1f %% busy (16 cycles each)\n", K,
           warps_per_sched, per, 100.0 * 16.0 * warps_per_sched / per);
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    for (int w : {1, 2, 4}) {
        run<1>(w, out, cyc); run<2>(w, out, cyc); run<3>(w, out, cyc); run<4>(w, out, cyc); run<6>(w, out, cyc);
        run<8>(w, out, cyc); run<12>(w, out, cyc); run<16>(w, out, cyc); run<32>(w, out, cyc);
    }
    return 0;
}
