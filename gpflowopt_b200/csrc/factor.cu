// Once-per-BO-iteration state of one GP (off the hot path), all FP64 on the device:
//   K = k(X, X) + noise I ;  L = chol(K) ;  Linv = L^-1 ;  V = Linv Y ;  block-stream packing of Linv.
// Replaces the candidate-independent half of GPflow 0.5.0 GPR.build_predict (called from
// gpflowopt/scaling.py:189), which the reference re-runs inside every session.run.
//
// Simple blocked (64-wide) right-looking algorithms built from three kernels each:
//   Cholesky : potrf on the diagonal block -> row-wise triangular solve of the panel -> rank-64 update
//   inverse  : forward substitution applied to the identity, right-looking (row scale -> rank-64 update)
#include "gpfo_internal.cuh"

#define FB 64            // factor block size

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double kern_from_r2(int kind, double variance, double r2) {
    if (kind == GPFO_KERN_RBF) return variance * exp(-r2 / 2.0);
    const double r = sqrt(r2 + 1e-12);
    if (kind == GPFO_KERN_MATERN52) {
        const double s5 = 2.2360679774997898;   // np.sqrt(5.)
        return variance * (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * exp(-s5 * r);
    }
    const double s3 = 1.7320508075688772;       // np.sqrt(3.)
    return variance * (1.0 + s3 * r) * exp(-s3 * r);
}

// Xs = X / lengthscales (row-wise), xn = row sums of squares; rows >= M are zero.
__global__ void prepare_inputs_kernel(const double* __restrict__ X, int M, int Mp, int D,
                                      const double* __restrict__ ell, double* __restrict__ Xs,
                                      double* __restrict__ xn) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Mp) return;
    double s = 0.0;
    for (int d = 0; d < D; ++d) {
        double v = (i < M) ? X[(size_t)i * D + d] / ell[d] : 0.0;
        Xs[(size_t)i * D + d] = v;
        s += v * v;
    }
    xn[i] = s;
}

void gpfo_prepare_inputs(cudaStream_t st, const double* dX, int M, int Mp, int D, const double* d_ell,
                         double* dXs, double* dxn) {
    prepare_inputs_kernel<<<(Mp + 127) / 128, 128, 0, st>>>(dX, M, Mp, D, d_ell, dXs, dxn);
}

// K (lower triangle and diagonal; the strict upper triangle is zeroed), identity on the padding.
__global__ void build_K_kernel(const double* __restrict__ Xs, const double* __restrict__ xn, int M, int Mf,
                               int D, int kind, double variance, double noise, double* __restrict__ K) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= Mf || j >= Mf) return;
    double v = 0.0;
    if (i >= M || j >= M) v = (i == j) ? 1.0 : 0.0;
    else if (j <= i) {
        double dot = 0.0;
        for (int d = 0; d < D; ++d) dot += Xs[(size_t)i * D + d] * Xs[(size_t)j * D + d];
        double r2 = (-2.0 * dot + xn[i]) + xn[j];            // GPflow square_dist, expanded form
        v = kern_from_r2(kind, variance, r2);
        if (i == j) v += noise;                               // + eye * likelihood.variance
    }
    K[(size_t)i * Mf + j] = v;
}

// Unblocked Cholesky of the FB x FB diagonal block k (in place, lower).  flag <- k+1 on a bad pivot.
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ A, int ld, int k, int* flag) {
    __shared__ double T[FB][FB + 1];
    __shared__ int bad;
    double* blk = A + (size_t)k * FB * ld + (size_t)k * FB;
    const int tid = threadIdx.x;
    if (tid == 0) bad = 0;
    for (int e = tid; e < FB * FB; e += 256) T[e / FB][e % FB] = blk[(size_t)(e / FB) * ld + (e % FB)];
    __syncthreads();
    for (int c = 0; c < FB; ++c) {
        if (tid == 0) {
            double d = T[c][c];
            if (!(d > 0.0)) { bad = 1; d = 1.0; }
            T[c][c] = sqrt(d);
        }
        __syncthreads();
        const double piv = T[c][c];
        for (int i = c + 1 + tid; i < FB; i += 256) T[i][c] = T[i][c] / piv;
        __syncthreads();
        const int n = FB - c - 1;
        for (int e = tid; e < n * n; e += 256) {
            int i = c + 1 + e / n, j = c + 1 + e % n;
            if (j <= i) T[i][j] -= T[i][c] * T[j][c];
        }
        __syncthreads();
    }
    for (int e = tid; e < FB * FB; e += 256) {
        int i = e / FB, j = e % FB;
        blk[(size_t)i * ld + j] = (j <= i) ? T[i][j] : 0.0;
    }
    if (tid == 0 && bad) atomicCAS(flag, 0, k + 1);
}

// Panel: rows below block k, columns of block k:  X * Lkk^T = A  (one thread per row).
__global__ void __launch_bounds__(FB) trsm_panel_kernel(double* __restrict__ A, int ld, int k) {
    extern __shared__ double dyn_smem[];
    double (*Lk)[FB + 1] = reinterpret_cast<double (*)[FB + 1]>(dyn_smem);
    double (*T)[FB + 1] = reinterpret_cast<double (*)[FB + 1]>(dyn_smem + FB * (FB + 1));   // T[c][row]
    const int tid = threadIdx.x;
    const double* lkk = A + (size_t)k * FB * ld + (size_t)k * FB;
    double* rows = A + ((size_t)(k + 1 + blockIdx.x) * FB) * ld + (size_t)k * FB;
    for (int e = tid; e < FB * FB; e += FB) {
        Lk[e / FB][e % FB] = lkk[(size_t)(e / FB) * ld + (e % FB)];
        T[e % FB][e / FB] = rows[(size_t)(e / FB) * ld + (e % FB)];
    }
    __syncthreads();
    for (int c = 0; c < FB; ++c) {
        double s = T[c][tid];
        for (int t = 0; t < c; ++t) s -= T[t][tid] * Lk[c][t];
        T[c][tid] = s / Lk[c][c];
    }
    __syncthreads();
    for (int e = tid; e < FB * FB; e += FB) rows[(size_t)(e / FB) * ld + (e % FB)] = T[e % FB][e / FB];
}

// C[64x64 tile] -= A[64 x 64] * op(B).  NT: op(B)[t][j] = B[j][t];  NN: op(B)[t][j] = B[t][j].
// 256 threads, 4x4 outputs each, K staged in two halves of 32.
template <bool NT>
__device__ __forceinline__ void rank64_tile(const double* __restrict__ Ag, const double* __restrict__ Bg,
                                            double* __restrict__ Cg, int ld) {
    __shared__ double As[FB][33];
    __shared__ double Bs[NT ? FB : 32][NT ? 33 : FB + 1];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int h = 0; h < 2; ++h) {
        for (int e = tid; e < FB * 32; e += 256) {
            int r = e / 32, c = e % 32;
            As[r][c] = Ag[(size_t)r * ld + h * 32 + c];
        }
        if (NT) {
            for (int e = tid; e < FB * 32; e += 256) {
                int r = e / 32, c = e % 32;
                Bs[r][c] = Bg[(size_t)r * ld + h * 32 + c];
            }
        } else {
            for (int e = tid; e < 32 * FB; e += 256) {
                int r = e / FB, c = e % FB;
                Bs[r][c] = Bg[(size_t)(h * 32 + r) * ld + c];
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int t = 0; t < 32; ++t) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = As[ty + 16 * a][t];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = NT ? Bs[tx + 16 * b][t] : Bs[t][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] += av[a] * bv[b];
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) Cg[(size_t)(ty + 16 * a) * ld + tx + 16 * b] -= acc[a][b];
}

// Cholesky trailing update after panel k: A[ti][tj] -= P[ti] P[tj]^T for ti >= tj > k.
__global__ void __launch_bounds__(256) syrk_update_kernel(double* __restrict__ A, int ld, int k) {
    const int tj = k + 1 + blockIdx.x, ti = k + 1 + blockIdx.y;
    if (tj > ti) return;
    const double* Pi = A + (size_t)ti * FB * ld + (size_t)k * FB;
    const double* Pj = A + (size_t)tj * FB * ld + (size_t)k * FB;
    rank64_tile<true>(Pi, Pj, A + (size_t)ti * FB * ld + (size_t)tj * FB, ld);
}

// Inverse of every FB x FB diagonal block of L (one CTA per block, one thread per column).
__global__ void __launch_bounds__(FB) trtri_diag_kernel(const double* __restrict__ L, int ld,
                                                        double* __restrict__ Dinv) {
    extern __shared__ double dyn_smem[];
    double (*Lk)[FB + 1] = reinterpret_cast<double (*)[FB + 1]>(dyn_smem);
    double (*Xc)[FB + 1] = reinterpret_cast<double (*)[FB + 1]>(dyn_smem + FB * (FB + 1));  // Xc[i][col]
    const int tid = threadIdx.x, k = blockIdx.x;
    const double* lkk = L + (size_t)k * FB * ld + (size_t)k * FB;
    for (int e = tid; e < FB * FB; e += FB) Lk[e / FB][e % FB] = lkk[(size_t)(e / FB) * ld + (e % FB)];
    __syncthreads();
    for (int i = 0; i < FB; ++i) {
        double s = (i == tid) ? 1.0 : 0.0;
        if (i >= tid) {
            for (int t = tid; t < i; ++t) s -= Lk[i][t] * Xc[t][tid];
            s = s / Lk[i][i];
        }
        Xc[i][tid] = (i >= tid) ? s : 0.0;
    }
    __syncthreads();
    double* out = Dinv + (size_t)k * FB * FB;
    for (int e = tid; e < FB * FB; e += FB) out[e] = Xc[e / FB][e % FB];
}

// Inverse, step k, row scale: X[k][j] = Dinv_k * R[k][j] for tiles j <= k (R is stored in X).
__global__ void __launch_bounds__(256) inv_rowscale_kernel(double* __restrict__ X, int ld,
                                                           const double* __restrict__ Dinv, int k) {
    extern __shared__ double dyn_smem[];
    double (*Ds)[FB + 1] = reinterpret_cast<double (*)[FB + 1]>(dyn_smem);
    double (*Rs)[FB + 1] = reinterpret_cast<double (*)[FB + 1]>(dyn_smem + FB * (FB + 1));
    const int tid = threadIdx.x, j = blockIdx.x;
    const double* dk = Dinv + (size_t)k * FB * FB;
    double* tile = X + (size_t)k * FB * ld + (size_t)j * FB;
    for (int e = tid; e < FB * FB; e += 256) {
        Ds[e / FB][e % FB] = dk[e];
        Rs[e / FB][e % FB] = tile[(size_t)(e / FB) * ld + (e % FB)];
    }
    __syncthreads();
    for (int e = tid; e < FB * FB; e += 256) {
        int r = e / FB, c = e % FB;
        double s = 0.0;
        for (int t = 0; t <= r; ++t) s += Ds[r][t] * Rs[t][c];      // Dinv_k is lower triangular
        tile[(size_t)r * ld + c] = s;
    }
}

// Inverse, step k, update: R[i][j] -= L[i][k] * X[k][j] for i > k, j <= k.
__global__ void __launch_bounds__(256) inv_update_kernel(double* __restrict__ X, const double* __restrict__ L,
                                                         int ld, int k) {
    const int j = blockIdx.x, i = k + 1 + blockIdx.y;
    rank64_tile<false>(L + (size_t)i * FB * ld + (size_t)k * FB, X + (size_t)k * FB * ld + (size_t)j * FB,
                       X + (size_t)i * FB * ld + (size_t)j * FB, ld);
}

__global__ void set_identity_kernel(double* __restrict__ X, int n) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    X[e] = (e / n == e % n) ? 1.0 : 0.0;
}

// V[i][r] = sum_{j<=i} Linv[i][j] Y[j][r]   (one warp per row, fixed lane-strided order + shuffle tree)
__global__ void linv_times_y_kernel(const double* __restrict__ Linv, int ld, const double* __restrict__ Y, int Mp,
                                    int R, double* __restrict__ V) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (warp >= Mp) return;
    for (int r = 0; r < R; ++r) {
        double s = 0.0;
        for (int j = lane; j <= warp; j += 32) s += Linv[(size_t)warp * ld + j] * Y[(size_t)j * R + r];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) V[(size_t)warp * R + r] = s;
    }
}

// One thread per packed double (see PackGeom in gpfo_internal.cuh).
__global__ void pack_linv_kernel(const double* __restrict__ Linv, int ld, PackGeom g, int p, int nr,
                                 int64_t chunk_base, double* __restrict__ packed) {
    // grid.x enumerates (kp, rb_local) pairs of chunk p densely over [0, kp_end) x [0, nr); inactive pairs exit
    const int kp_end = g.dense ? g.nrb : p * g.rbc + nr;
    const int64_t pair = blockIdx.x;
    const int kp = (int)(pair / nr), rbl = (int)(pair % nr);
    if (kp >= kp_end) return;
    int lo;
    const int64_t cb = pg_col_base(g, p, nr, kp, &lo);
    if (rbl < lo) return;
    const int e = threadIdx.x;                 // 0..63
    const int lane = e >> 1, h = e & 1;
    const int row = 8 * (p * g.rbc + rbl) + (lane >> 2);
    const int col = 8 * kp + 4 * h + (lane & 3);
    packed[(chunk_base + cb + (rbl - lo)) * 64 + e] = Linv[(size_t)row * ld + col];
}

void gpfo_pack_linv(cudaStream_t st, const double* dLinv, int ld, PackGeom g, double* dpacked) {
    for (int p = 0; p < g.nchunks; ++p) {
        const int nr = pg_chunk_rows(g, p);
        const int kp_end = g.dense ? g.nrb : p * g.rbc + nr;
        const int64_t pairs = (int64_t)kp_end * nr;
        pack_linv_kernel<<<(unsigned)pairs, 64, 0, st>>>(dLinv, ld, g, p, nr, pg_chunk_base(g, p), dpacked);
    }
}

int gpfo_factor_device(cudaStream_t st, int M, int Mf, int D, int R, int kind, double variance, double noise,
                       const double* dXs, const double* dxn, double* dK, double* dLinv, const double* dY,
                       double* dV, int* dflag, std::string* err) {
    // dK, dLinv: Mf x Mf with Mf a multiple of FB; dXs/dxn/dY/dV have at least Mf rows (zero padded).
    const int nb = Mf / FB;
    const int two_tiles = 2 * FB * (FB + 1) * (int)sizeof(double);
    cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, two_tiles);
    cudaFuncSetAttribute(trtri_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, two_tiles);
    cudaFuncSetAttribute(inv_rowscale_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, two_tiles);
    cudaMemsetAsync(dflag, 0, sizeof(int), st);
    {
        dim3 blk(32, 8), grd((Mf + 31) / 32, (Mf + 7) / 8);
        build_K_kernel<<<grd, blk, 0, st>>>(dXs, dxn, M, Mf, D, kind, variance, noise, dK);
    }
    for (int k = 0; k < nb; ++k) {
        potrf_diag_kernel<<<1, 256, 0, st>>>(dK, Mf, k, dflag);
        const int rem = nb - k - 1;
        if (rem > 0) {
            trsm_panel_kernel<<<rem, FB, two_tiles, st>>>(dK, Mf, k);
            syrk_update_kernel<<<dim3(rem, rem), 256, 0, st>>>(dK, Mf, k);
        }
    }
    int flag = 0;
    cudaError_t ce = cudaMemcpyAsync(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { *err = std::string("factor: ") + cudaGetErrorString(ce); return GPFO_ERR_CUDA; }
    if (flag != 0) {
        *err = "Cholesky decomposition was not successful: non-positive pivot in block " + std::to_string(flag - 1) +
               " (K + noise*I is not positive definite)";
        return GPFO_ERR_NOT_SPD;
    }
    // inverse: the diagonal-block inverses live in the (unused) strict upper part?  no -- separate scratch at the
    // tail of dLinv is not available either, so allocate a small buffer here (nb * 32 KB).
    double* dDinv = nullptr;
    ce = cudaMallocAsync((void**)&dDinv, (size_t)nb * FB * FB * sizeof(double), st);
    if (ce != cudaSuccess) { *err = std::string("factor: ") + cudaGetErrorString(ce); return GPFO_ERR_CUDA; }
    trtri_diag_kernel<<<nb, FB, two_tiles, st>>>(dK, Mf, dDinv);
    {
        size_t n2 = (size_t)Mf * Mf;
        set_identity_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(dLinv, Mf);
    }
    for (int k = 0; k < nb; ++k) {
        inv_rowscale_kernel<<<k + 1, 256, two_tiles, st>>>(dLinv, Mf, dDinv, k);
        const int rem = nb - k - 1;
        if (rem > 0) inv_update_kernel<<<dim3(k + 1, rem), 256, 0, st>>>(dLinv, dK, Mf, k);
    }
    cudaFreeAsync(dDinv, st);
    {
        const int warps_per_block = 4;
        linv_times_y_kernel<<<(Mf + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(
            dLinv, Mf, dY, Mf, R, dV);
    }
    ce = cudaGetLastError();
    if (ce != cudaSuccess) { *err = std::string("factor: ") + cudaGetErrorString(ce); return GPFO_ERR_CUDA; }
    return GPFO_OK;
}

// Kinv[i][j] = sum_t Linv[t][i] * Linv[t][j]  (t >= max(i, j)); 64x64 tile per CTA, 256 threads, 4x4 outputs each.
__global__ void __launch_bounds__(256) ata_kernel(const double* __restrict__ A, int ld, double* __restrict__ C) {
    __shared__ double As[32][FB + 1];
    __shared__ double Bs[32][FB + 1];
    const int ti = blockIdx.y, tj = blockIdx.x;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const int t0 = (ti > tj ? ti : tj) * FB;              // rows above the larger tile index are zero in both
    for (int t = t0; t < ld; t += 32) {
        for (int e = tid; e < 32 * FB; e += 256) {
            const int r = e / FB, c = e % FB;
            As[r][c] = A[(size_t)(t + r) * ld + ti * FB + c];
            Bs[r][c] = A[(size_t)(t + r) * ld + tj * FB + c];
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = As[k][ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = Bs[k][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] += av[a] * bv[b];
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) C[(size_t)(ti * FB + ty + 16 * a) * ld + tj * FB + tx + 16 * b] = acc[a][b];
}

// alpha[i][r] = sum_{t >= i} Linv[t][i] V[t][r]   (one warp per column i)
__global__ void linvT_times_v_kernel(const double* __restrict__ Linv, int ld, const double* __restrict__ V, int R,
                                     double* __restrict__ alpha) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (i >= ld) return;
    for (int r = 0; r < R; ++r) {
        double s = 0.0;
        for (int t = i + lane; t < ld; t += 32) s += Linv[(size_t)t * ld + i] * V[(size_t)t * R + r];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) alpha[(size_t)i * R + r] = s;
    }
}

void gpfo_kinv_alpha(cudaStream_t st, const double* dLinv, int Mf, const double* dV, int R, double* dKinv, double* dalpha) {
    const int nb = Mf / FB;
    ata_kernel<<<dim3(nb, nb), 256, 0, st>>>(dLinv, Mf, dKinv);
    linvT_times_v_kernel<<<(Mf + 3) / 4, 128, 0, st>>>(dLinv, Mf, dV, R, dalpha);
}

// dst[i][k] = src[n-1-k][n-1-i]: the transpose of a lower-triangular matrix, re-indexed so that it is lower triangular
// again (stored-A backward pass: W = L^-T A runs through the same triangular block-stream code as A = L^-1 Kx)
__global__ void reverse_transpose_kernel(const double* __restrict__ src, int ld, int n, double* __restrict__ dst) {
    __shared__ double t[32][33];
    const int i0 = blockIdx.y * 32, k0 = blockIdx.x * 32;       // dst tile rows i0.., cols k0..
    for (int r = threadIdx.y; r < 32; r += 8) {                  // read src rows n-1-(k0+r), cols n-1-(i0+c): coalesced over c
        const int k = k0 + r, i = i0 + threadIdx.x;
        t[r][threadIdx.x] = (k < n && i < n) ? src[(size_t)(n - 1 - k) * ld + (n - 1 - i)] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = i0 + r, k = k0 + threadIdx.x;
        if (i < n && k < n) dst[(size_t)i * ld + k] = t[threadIdx.x][r];
    }
}

void gpfo_reverse_transpose(cudaStream_t st, const double* src, int ld, int n, double* dst) {
    dim3 blk(32, 8), grd((n + 31) / 32, (n + 31) / 32);
    reverse_transpose_kernel<<<grd, blk, 0, st>>>(src, ld, n, dst);
}

#ifdef GPFO_BUILD_EXPERIMENTS
// ---- cluster layout (see ClusterGeom) -----------------------------------------------------------------------------
void gpfo_cluster_geom(int nrb, int rbc, ClusterGeom* g, long long total[2]) {
    g->nrb = nrb;
    g->rbc = rbc;
    g->nchunks = (nrb + 2 * rbc - 1) / (2 * rbc);
    for (int c = 0; c < 2; ++c) {
        long long off = 0;
        for (int p = 0; p < g->nchunks; ++p) {
            g->base[c][p] = off;
            const int nr = cg_rows(*g, p, c), kd0 = p * 2 * rbc, ke = cg_kp_end(*g, p);
            for (int kp = 0; kp < ke; ++kp) off += nr - cg_lo(kp - kd0, c);
        }
        total[c] = off;
    }
}

__global__ void pack_linv_cluster_kernel(const double* __restrict__ Linv, int ld, ClusterGeom g, int p, int c, int nr,
                                         double* __restrict__ packed) {
    const int kd0 = p * 2 * g.rbc;
    const int kp = blockIdx.x / nr, q = blockIdx.x % nr;
    const int lo = cg_lo(kp - kd0, c);
    if (q < lo) return;
    long long off = g.base[c][p];
    if (kp <= kd0) off += (long long)kp * nr;
    else {
        off += (long long)kd0 * nr;
        for (int k = kd0; k < kp; ++k) off += nr - cg_lo(k - kd0, c);
    }
    const int e = threadIdx.x, lane = e >> 1, h = e & 1;
    const int row = 8 * (kd0 + 2 * q + c) + (lane >> 2);
    const int col = 8 * kp + 4 * h + (lane & 3);
    packed[(off + (q - lo)) * 64 + e] = Linv[(size_t)row * ld + col];
}

void gpfo_pack_linv_cluster(cudaStream_t st, const double* dLinv, int ld, const ClusterGeom& g, double* dpacked0,
                            double* dpacked1) {
    for (int c = 0; c < 2; ++c)
        for (int p = 0; p < g.nchunks; ++p) {
            const int nr = cg_rows(g, p, c);
            if (nr == 0) continue;
            const long long pairs = (long long)cg_kp_end(g, p) * nr;
            pack_linv_cluster_kernel<<<(unsigned)pairs, 64, 0, st>>>(dLinv, ld, g, p, c, nr, c ? dpacked1 : dpacked0);
        }
}
#endif  // GPFO_BUILD_EXPERIMENTS
