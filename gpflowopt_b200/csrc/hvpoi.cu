// Second kernel of the path: per-candidate sum over Pareto cells of Gaussian CDF products (HVPoI), followed by
// the rest of the acquisition program and the block argmax.
//
// Replaces gpflowopt/acquisition/hvpoi.py:104-140: the Phi table normal.cdf(pf_ext) (N x (P+2) x R), the two
// gather_nd's over bounds.ub / bounds.lb (N x C x R temporaries in the reference), PoI = sum_c prod_j dPhi and
// Hv = sum_c [all_j ub > mu] prod_j (ub - max(lb, mu)).  Here one CTA owns TC candidates, keeps their Phi
// tables and pf_ext in shared memory, and streams the int32 cell tables once per TC candidates.
#include <cstdlib>
#include "gpfo_internal.cuh"
#include "ndtr_fast.cuh"

#define HV_TC 8            // candidates per CTA tile
#define HV_THREADS 256
#define HV_RMAX 8          // objectives

// Shared-memory tables are laid out [pf_ext row][objective][candidate] with HV_TC = 8 candidates innermost and every row
// padded by 16 bytes (HV_ROW(R) doubles): a thread fetches the 8 candidates of one (row, objective) entry with four
// 16-byte loads, and rows that differ by 1..7 fall into disjoint bank groups (row stride = odd multiple of 16 B mod 128),
// so the per-cell gathers of a quarter warp are conflict free.  (First version: [candidate][row][objective] with scalar
// gathers -> 143 ms for config #4; see profiles/r01_sweep_configs2-5.log.)
#define HV_ROW(R) ((R) * HV_TC + 2)

// GRAD: also d(Hv PoI)/d mean_j and /d var_j per candidate (closed form of SURVEY Appendix A item 10), fed through the
// forward-mode program evaluation; writes the per-slot coefficients the contraction kernel's gradient mode consumes.
template <int R, bool GRAD>
__global__ void __launch_bounds__(HV_THREADS, GRAD ? 1 : 2)
hvpoi_kernel(const ProgramDev* __restrict__ prog, CellsDev cells, EvalIO io, int slot_first) {
    extern __shared__ __align__(16) double sm[];
    const int n_ext = cells.n_ext;
    constexpr int ROW = HV_ROW(R);
    constexpr int NQ = GRAD ? 2 + 3 * R : 2;           // per-candidate sums: PoI, Hv [, dPoI/dmu_j, dPoI/dsd_j, dHv/dmu_j]
    double* phi = sm;                                  // [n_ext][ROW]   Phi(z)
    double* pdf = phi + (size_t)n_ext * ROW;           // GRAD: [n_ext][ROW] phi(z)
    double* zpdf = pdf + (GRAD ? (size_t)n_ext * ROW : 0);           // GRAD: [n_ext][ROW] z phi(z)
    double* pf = zpdf + (GRAD ? (size_t)n_ext * ROW : 0);            // [n_ext][R]
    double* cmu = pf + (((size_t)n_ext * R + 1) & ~(size_t)1);    // [R][HV_TC]  (kept 16-byte aligned)
    double* csd = cmu + HV_TC * R;                     // [R][HV_TC]
    double* red = csd + HV_TC * R;                     // [HV_THREADS/32][HV_TC][NQ]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    for (int e = tid; e < n_ext * R; e += HV_THREADS) pf[e] = cells.pf_ext[e];

    double best_v = 0.0;
    long long best_i = -1;
    const int64_t ntiles = (io.N + HV_TC - 1) / HV_TC;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * HV_TC;
        __syncthreads();
        if (tid < HV_TC * R) {
            const int j = tid / HV_TC, t = tid % HV_TC;
            const int64_t n = n0 + t;
            double m = 0.0, v = 1.0;
            if (n < io.N) {
                m = io.mom_mean[(size_t)(slot_first + j) * io.mstride + n];
                v = io.mom_var[(size_t)(slot_first + j) * io.mstride + n];
            }
            v = fmax(v, GPFO_STABILITY);               // hvpoi.py:109
            cmu[tid] = m;
            csd[tid] = sqrt(v);
        }
        __syncthreads();
        // Phi[i][j][t] = Normal(mu_tj, sd_tj).cdf(pf_ext[i][j])                           (hvpoi.py:112-113)
        for (int e = tid; e < n_ext * R * HV_TC; e += HV_THREADS) {
            const int i = e / (R * HV_TC), jt = e % (R * HV_TC), j = jt / HV_TC;
            const double z = (pf[i * R + j] - cmu[jt]) / csd[jt];
            phi[i * ROW + jt] = gpfo_ndtr(z);
            if (GRAD) {                                // terms with infinite z vanish
                const bool fin = isfinite(z);
                const double pz = fin ? gpfo_std_pdf(z) : 0.0;
                pdf[i * ROW + jt] = pz;
                zpdf[i * ROW + jt] = fin ? z * pz : 0.0;
            }
        }
        __syncthreads();
        double q[HV_TC][NQ];
#pragma unroll
        for (int t = 0; t < HV_TC; ++t)
#pragma unroll
            for (int k = 0; k < NQ; ++k) q[t][k] = 0.0;
        for (int64_t c = tid; c < cells.C; c += HV_THREADS) {
            int ub[R], lb[R];
            double ubp[R], lbp[R];
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int iu = cells.ub[c * R + j], il = cells.lb[c * R + j];
                ubp[j] = pf[iu * R + j];
                lbp[j] = pf[il * R + j];
                ub[j] = iu * ROW + j * HV_TC;
                lb[j] = il * ROW + j * HV_TC;
            }
            if (!GRAD) {
                double pr[HV_TC], ext[HV_TC];
                bool valid[HV_TC];
#pragma unroll
                for (int t = 0; t < HV_TC; ++t) { pr[t] = 1.0; ext[t] = 1.0; valid[t] = true; }
#pragma unroll
                for (int j = 0; j < R; ++j) {
#pragma unroll
                    for (int t2 = 0; t2 < HV_TC / 2; ++t2) {
                        const double2 pu = *reinterpret_cast<const double2*>(phi + ub[j] + 2 * t2);
                        const double2 pl = *reinterpret_cast<const double2*>(phi + lb[j] + 2 * t2);
                        const double2 m2 = *reinterpret_cast<const double2*>(cmu + j * HV_TC + 2 * t2);
                        pr[2 * t2] *= (pu.x - pl.x);                                        // hvpoi.py:116-120
                        pr[2 * t2 + 1] *= (pu.y - pl.y);
                        valid[2 * t2] = valid[2 * t2] && (ubp[j] > m2.x);                   // hvpoi.py:130-131
                        valid[2 * t2 + 1] = valid[2 * t2 + 1] && (ubp[j] > m2.y);
                        ext[2 * t2] *= (ubp[j] - fmax(lbp[j], m2.x));                       // hvpoi.py:133-136
                        ext[2 * t2 + 1] *= (ubp[j] - fmax(lbp[j], m2.y));
                    }
                }
#pragma unroll
                for (int t = 0; t < HV_TC; ++t) {
                    q[t][0] += pr[t];
                    q[t][1] += valid[t] ? ext[t] : 0.0;
                }
            } else {
#pragma unroll
                for (int t = 0; t < HV_TC; ++t) {
                    double pr = 1.0, ext = 1.0;
                    bool valid = true;
                    double dP[R], ex[R];
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        dP[j] = phi[ub[j] + t] - phi[lb[j] + t];
                        pr *= dP[j];
                        const double m = cmu[j * HV_TC + t];
                        valid = valid && (ubp[j] > m);
                        ex[j] = ubp[j] - fmax(lbp[j], m);
                        ext *= ex[j];
                    }
                    q[t][0] += pr;
                    q[t][1] += valid ? ext : 0.0;
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        double oP = 1.0, oE = 1.0;                                         // exclusive products
#pragma unroll
                        for (int k = 0; k < R; ++k) if (k != j) { oP *= dP[k]; oE *= ex[k]; }
                        const double isd = 1.0 / csd[j * HV_TC + t];
                        q[t][2 + j] += oP * (-(pdf[ub[j] + t] - pdf[lb[j] + t]) * isd);    // d PoI / d mu_j
                        q[t][2 + R + j] += oP * (-(zpdf[ub[j] + t] - zpdf[lb[j] + t]) * isd);   // d PoI / d sd_j
                        // tf.maximum(lb, mu): the gradient reaches mu only where lb < mu; the indicator carries none
                        q[t][2 + 2 * R + j] += (valid && lbp[j] < cmu[j * HV_TC + t]) ? -oE : 0.0;   // d Hv / d mu_j
                    }
                }
            }
        }
        // deterministic block reduction: shuffle tree inside the warp, fixed warp order across warps
#pragma unroll
        for (int t = 0; t < HV_TC; ++t)
#pragma unroll
            for (int k = 0; k < NQ; ++k) {
                double a = q[t][k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) red[(w * HV_TC + t) * NQ + k] = a;
            }
        __syncthreads();
        if (w == 0) {
            double val = 0.0;
            long long idx = -1;
            const int64_t n = n0 + lane;
            if (lane < HV_TC && n < io.N) {
                double tq[NQ];
#pragma unroll
                for (int k = 0; k < NQ; ++k) {
                    double a = 0.0;
                    for (int ww = 0; ww < HV_THREADS / 32; ++ww) a += red[(ww * HV_TC + lane) * NQ + k];
                    tq[k] = a;
                }
                const double leaf = tq[1] * tq[0];                                         // Hv * PoI, hvpoi.py:140
                double mu[GPFO_MAX_SLOTS], va[GPFO_MAX_SLOTS];
                for (int s = 0; s < prog->nslots; ++s) {
                    mu[s] = io.mom_mean[(size_t)s * io.mstride + n];
                    va[s] = io.mom_var[(size_t)s * io.mstride + n];
                }
                if (GRAD) {
                    double hdm[R], hdv[R], dmu[GPFO_MAX_SLOTS], dva[GPFO_MAX_SLOTS];
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        hdm[j] = tq[0] * tq[2 + 2 * R + j] + tq[1] * tq[2 + j];           // PoI dHv + Hv dPoI
                        const double v0 = va[slot_first + j];
                        const double ds_dv = (v0 >= GPFO_STABILITY) ? 1.0 / (2.0 * csd[j * HV_TC + lane]) : 0.0;
                        hdv[j] = tq[1] * tq[2 + R + j] * ds_dv;
                    }
                    val = gpfo_run_program<false>(*prog, mu, va, leaf, nullptr, nullptr, nullptr, nullptr);   // same code as evaluate
                    gpfo_run_program<true>(*prog, mu, va, leaf, hdm, hdv, dmu, dva);
                    for (int s = 0; s < prog->nslots; ++s) {
                        io.cmu[(size_t)s * io.mstride + n] = dmu[s];
                        io.cva[(size_t)s * io.mstride + n] = dva[s];
                    }
                } else
                val = gpfo_run_program<false>(*prog, mu, va, leaf, nullptr, nullptr, nullptr, nullptr);
                if (io.values) io.values[n] = val;
                idx = io.idx0 + n;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, val, o);
                const long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
                if (gpfo_better(ov, oi, val, idx)) { val = ov; idx = oi; }
            }
            if (gpfo_better(val, idx, best_v, best_i)) { best_v = val; best_i = idx; }
        }
    }
    if (tid == 0) {
        io.part_val[blockIdx.x] = best_v;
        io.part_idx[blockIdx.x] = best_i;
    }
}


// ------------------------------------------------------------------------------------------------------------------------
// hvpoi2_kernel (values only; the default when the Phi table fits): ONE CANDIDATE PER LANE.
//
// The first kernel above maps threads to cells and gathers the Phi entries of 8 candidates per cell: its int32 index math is
// per thread, 38 % of its shared-memory wavefronts are bank conflicts of those gathers, and its Phi table costs more than its
// cell sums (libm erf / erfc diverge).  Here a CTA owns TC = 32 (or 16) candidates; the Phi table is [entry][candidate]
// (entry = pf_ext row * R + objective), so for a given cell -- the same cell for all lanes of a warp (TC = 32) or of a
// half-warp (TC = 16) -- lane t reads phi[entry * TC + t]: one contiguous row, conflict free, and every index is warp-uniform.
// Warps split the cells; the per-cell records (entry offsets + the two bounds per objective) are precomputed once by
// gpfo_cells_set.  Phi uses the branch-free erfc of ndtr_fast.cuh.  Sums are taken in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------------------------------------
#define HV2_THREADS 512

template <int R, int TC>
__global__ void __launch_bounds__(HV2_THREADS, 1)
hvpoi2_kernel(const ProgramDev* __restrict__ prog, CellsDev cells, EvalIO io, int slot_first) {
    extern __shared__ __align__(16) double sm[];
    constexpr int NW = HV2_THREADS / 32;
    constexpr int CW = 32 / TC;                        // cells per warp iteration
    const int n_ent = cells.n_ext * R;
    double* phi = sm;                                  // [n_ent][TC]
    double* pf = phi + (size_t)n_ent * TC;             // [n_ent]
    double* cmu = pf + ((n_ent + 1) & ~1);             // [R][TC]
    double* cis = cmu + R * TC;                        // [R][TC]   1 / sd
    double* red = cis + R * TC;                        // [NW][TC][2]
    double* tab32 = red + NW * TC * 2;                 // [32]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int t = lane & (TC - 1), sub = lane / TC;    // candidate of this lane, cell slot within the warp iteration

    for (int e = tid; e < n_ent; e += HV2_THREADS) pf[e] = cells.pf_ext[e];
    if (tid < 32) tab32[tid] = exp2(tid * (1.0 / 32.0));

    double best_v = 0.0;
    long long best_i = -1;
    const int64_t ntiles = (io.N + TC - 1) / TC;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = tile * TC;
        __syncthreads();
        if (tid < TC * R) {
            const int j = tid / TC, tt = tid % TC;
            const int64_t n = n0 + tt;
            double m = 0.0, v = 1.0;
            if (n < io.N) {
                m = io.mom_mean[(size_t)(slot_first + j) * io.mstride + n];
                v = io.mom_var[(size_t)(slot_first + j) * io.mstride + n];
            }
            v = fmax(v, GPFO_STABILITY);               // hvpoi.py:109
            cmu[tid] = m;
            cis[tid] = 1.0 / sqrt(v);
        }
        __syncthreads();
        // Phi[entry][t] = Normal(mu_tj, sd_tj).cdf(pf_ext[i][j])                            (hvpoi.py:112-113)
        for (int e = tid; e < n_ent * TC; e += HV2_THREADS) {
            const int ent = e / TC, tt = e - ent * TC;
            const int j = ent % R;
            phi[e] = gpfo_ndtr_fast((pf[ent] - cmu[j * TC + tt]) * cis[j * TC + tt], tab32);
        }
        __syncthreads();
        double mu[R];
#pragma unroll
        for (int j = 0; j < R; ++j) mu[j] = cmu[j * TC + t];
        double q0 = 0.0, q1 = 0.0;
        const double* __restrict__ phit = phi + t;
        // warp w, slot sub: cells (it * NW + w) * CW + sub
        for (int64_t c = (int64_t)w * CW + sub; c < cells.C; c += NW * CW) {
            const int2* __restrict__ off = reinterpret_cast<const int2*>(cells.off) + c * R;
            const double2* __restrict__ bnd = reinterpret_cast<const double2*>(cells.bnd) + c * R;
            double pr = 1.0, ext = 1.0;
            bool valid = true;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int2 o = __ldg(off + j);                                              // entry offsets of ub, lb
                const double2 b = __ldg(bnd + j);                                           // pf_ext values of ub, lb
                pr *= (phit[o.x * TC] - phit[o.y * TC]);                                    // hvpoi.py:116-120
                valid = valid && (b.x > mu[j]);                                             // hvpoi.py:130-131
                ext *= (b.x - fmax(b.y, mu[j]));                                            // hvpoi.py:133-136
            }
            q0 += pr;
            q1 += valid ? ext : 0.0;
        }
        if (CW == 2) {                                   // the two half-warps hold different cells of the same candidates
            q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
            q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
        }
        if (lane < TC) { red[(w * TC + t) * 2] = q0; red[(w * TC + t) * 2 + 1] = q1; }
        __syncthreads();
        if (w == 0) {
            double val = 0.0;
            long long idx = -1;
            const int64_t n = n0 + lane;
            if (lane < TC && n < io.N) {
                double poi = 0.0, hv = 0.0;
                for (int ww = 0; ww < NW; ++ww) { poi += red[(ww * TC + lane) * 2]; hv += red[(ww * TC + lane) * 2 + 1]; }
                const double leaf = hv * poi;                                               // Hv * PoI, hvpoi.py:140
                double m2[GPFO_MAX_SLOTS], va[GPFO_MAX_SLOTS];
                for (int s = 0; s < prog->nslots; ++s) {
                    m2[s] = io.mom_mean[(size_t)s * io.mstride + n];
                    va[s] = io.mom_var[(size_t)s * io.mstride + n];
                }
                val = gpfo_run_program<false>(*prog, m2, va, leaf, nullptr, nullptr, nullptr, nullptr);
                if (io.values) io.values[n] = val;
                idx = io.idx0 + n;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, val, o);
                const long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
                if (gpfo_better(ov, oi, val, idx)) { val = ov; idx = oi; }
            }
            if (gpfo_better(val, idx, best_v, best_i)) { best_v = val; best_i = idx; }
        }
    }
    if (tid == 0) {
        io.part_val[blockIdx.x] = best_v;
        io.part_idx[blockIdx.x] = best_i;
    }
}

static size_t hv2_smem(int n_ext, int R, int TC) {
    const size_t n_ent = (size_t)n_ext * R;
    return (n_ent * TC + ((n_ent + 1) & ~(size_t)1) + 2 * (size_t)R * TC + (HV2_THREADS / 32) * TC * 2 + 32) * sizeof(double);
}
// candidates per CTA of hvpoi2_kernel for this front size: 32, 16, or 0 = table too large (the first kernel takes over)
static int hv2_tile(int n_ext, int R) {
    if (hv2_smem(n_ext, R, 32) <= (size_t)200 * 1024) return 32;
    if (hv2_smem(n_ext, R, 16) <= (size_t)200 * 1024) return 16;
    return 0;
}

template <int R, int TC>
static cudaError_t launch_hv2(cudaStream_t st, const ProgramDev* d_prog, const CellsDev& cells, EvalIO io, int grid, int slot_first) {
    const size_t smem = hv2_smem(cells.n_ext, R, TC);
    auto kern = hvpoi2_kernel<R, TC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, HV2_THREADS, smem, st>>>(d_prog, cells, io, slot_first);
    return cudaGetLastError();
}

static size_t hv_smem(int n_ext, int R, bool grad) {
    const int nq = grad ? 2 + 3 * R : 2;
    return ((size_t)n_ext * HV_ROW(R) * (grad ? 3 : 1) + (((size_t)n_ext * R + 1) & ~(size_t)1) + 2 * HV_TC * R +
            (HV_THREADS / 32) * HV_TC * nq) * sizeof(double);
}

template <int R, bool GRAD>
static cudaError_t launch_hv(cudaStream_t st, const ProgramDev* d_prog, const CellsDev& cells, EvalIO io, int grid,
                             int slot_first) {
    const size_t smem = hv_smem(cells.n_ext, R, GRAD);
    auto kern = hvpoi_kernel<R, GRAD>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, HV_THREADS, smem, st>>>(d_prog, cells, io, slot_first);
    return cudaGetLastError();
}

// with_grad < 0: upper bound over both kernels (size of the partial-argmax arrays)
int gpfo_hvpoi_grid(int64_t N, int num_sms, int n_ext, int R, int with_grad) {
    const int tc2 = (with_grad == 0 && n_ext > 0 && !getenv("GPFO_HVPOI_V1")) ? hv2_tile(n_ext, R) : 0;
    const int tc = tc2 ? tc2 : HV_TC;
    const int64_t ntiles = (N + tc - 1) / tc;
    const int64_t cap = (int64_t)num_sms * (tc2 ? 1 : 2);
    const int64_t cap_any = (int64_t)num_sms * 2;
    if (with_grad < 0) { const int64_t t8 = (N + HV_TC - 1) / HV_TC; return (int)(t8 < cap_any ? (t8 > 0 ? t8 : 1) : cap_any); }
    return (int)(ntiles < cap ? (ntiles > 0 ? ntiles : 1) : cap);
}

size_t gpfo_hvpoi_smem_bytes(int n_ext, int R) { return hv_smem(n_ext, R, false); }
size_t gpfo_hvpoi_smem_bytes_grad(int n_ext, int R) { return hv_smem(n_ext, R, true); }

template <int R>
static cudaError_t launch_hv_r(cudaStream_t st, const ProgramDev* d_prog, const CellsDev& cells, EvalIO io, int grid,
                               int slot_first, int with_grad) {
    if (with_grad) return launch_hv<R, true>(st, d_prog, cells, io, grid, slot_first);
    const int tc2 = getenv("GPFO_HVPOI_V1") ? 0 : hv2_tile(cells.n_ext, R);       // GPFO_HVPOI_V1=1: the first kernel (comparisons)
    if (tc2 == 32) return launch_hv2<R, 32>(st, d_prog, cells, io, grid, slot_first);
    if (tc2 == 16) return launch_hv2<R, 16>(st, d_prog, cells, io, grid, slot_first);
    return launch_hv<R, false>(st, d_prog, cells, io, grid, slot_first);
}

cudaError_t gpfo_hvpoi_launch(cudaStream_t st, const ProgramDev* d_prog, const CellsDev& cells, EvalIO io, int grid,
                              int slot_first, int with_grad) {
    switch (cells.R) {
        case 2: return launch_hv_r<2>(st, d_prog, cells, io, grid, slot_first, with_grad);
        case 3: return launch_hv_r<3>(st, d_prog, cells, io, grid, slot_first, with_grad);
        case 4: return launch_hv_r<4>(st, d_prog, cells, io, grid, slot_first, with_grad);
        case 5: return launch_hv_r<5>(st, d_prog, cells, io, grid, slot_first, with_grad);
        case 6: return launch_hv_r<6>(st, d_prog, cells, io, grid, slot_first, with_grad);
        default: return cudaErrorInvalidValue;
    }
}
