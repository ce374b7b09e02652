// Internal declarations shared by the translation units of libgpfo.so (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/gpfo.h"

#define GPFO_HD __host__ __device__ __forceinline__

// ------------------------------------------------------------------------------------------------
// Packed L^-1 layout ("block stream") consumed by the contraction kernel.
//
// L^-1 (Mp x Mp, Mp = M rounded up to 8, zero padded) is cut into 8x8 blocks (rb = row block,
// kp = column block).  Only blocks with rb >= kp are stored.  Rows are grouped into chunks of RBC row
// blocks; inside chunk p the blocks are ordered [kp][rb] so that, for a fixed kp, the row blocks a CTA
// walks through are contiguous.  One block is 64 doubles in "DMMA lane order": lane l (0..31) owns
// {Linv[8rb + l/4][8kp + l%4], Linv[8rb + l/4][8kp + 4 + l%4]} = the A fragments of the two m8n8k4 steps
// of that block, stored as one 16-byte pair at block_base + 2*l.
// ------------------------------------------------------------------------------------------------
struct PackGeom {
    int nrb;        // number of 8-row blocks = Mp / 8
    int rbc;        // row blocks per chunk
    int nchunks;    // ceil(nrb / rbc)
    int dense;      // 0: lower block-triangular (L^-1); 1: every block stored (K^-1, gradient pass)
};

GPFO_HD int pg_chunk_rows(const PackGeom& g, int p) {            // row blocks in chunk p
    int left = g.nrb - p * g.rbc;
    return left < g.rbc ? left : g.rbc;
}
// number of blocks stored before chunk p
GPFO_HD int64_t pg_chunk_base(const PackGeom& g, int p) {
    // full chunks q < p each hold q*rbc*rbc (off-diagonal) + rbc*(rbc+1)/2 (diagonal) blocks
    int64_t rbc = g.rbc;
    int64_t pp = p;
    if (g.dense) return pp * rbc * g.nrb;
    return rbc * rbc * (pp * (pp - 1) / 2) + pp * (rbc * (rbc + 1) / 2);
}
// offset (in blocks, relative to the chunk base) of the first stored block of column block kp in chunk p,
// and the first stored row block (local index) `lo`.  nr = pg_chunk_rows(g, p).
GPFO_HD int64_t pg_col_base(const PackGeom& g, int p, int nr, int kp, int* lo) {
    int kd = kp - p * g.rbc;                                     // < 0: off-diagonal (dense) column block
    if (kd < 0 || g.dense) { *lo = 0; return (int64_t)kp * nr; }
    *lo = kd;
    // sum_{q<kd} (nr - q) = kd*nr - kd*(kd-1)/2
    return (int64_t)p * g.rbc * nr + (int64_t)kd * nr - (int64_t)kd * (kd - 1) / 2;
}
GPFO_HD int64_t pg_total_blocks(const PackGeom& g) {
    if (g.dense) return (int64_t)g.nrb * g.nrb;
    int p = g.nchunks - 1;
    int nr = pg_chunk_rows(g, p);
    return pg_chunk_base(g, p) + (int64_t)p * g.rbc * nr + (int64_t)nr * (nr + 1) / 2;
}

#ifdef GPFO_BUILD_EXPERIMENTS
// ------------------------------------------------------------------------------------------------
// Cluster layout (contract3.cu): a pair of CTAs (one thread-block cluster) shares one candidate tile.  Rows are cut
// into cluster chunks of 2*rbc row blocks; CTA rank c owns the row blocks of parity c (local block q <-> global block
// 2*rbc*p + 2q + c), so both CTAs walk the same column blocks in lock step and the triangular work is balanced.
// Each rank has its own block stream, ordered [chunk][kp][q] with q >= cg_lo(kd, c) only.
// ------------------------------------------------------------------------------------------------
#define GPFO_MAX_CCHUNKS 64
struct ClusterGeom {
    int nrb;        // number of 8-row blocks
    int rbc;        // row blocks per chunk PER CTA
    int nchunks;
    long long base[2][GPFO_MAX_CCHUNKS];     // first block of (rank, chunk) in that rank's stream
};
GPFO_HD int cg_rows(const ClusterGeom& g, int p, int c) {
    const int first = p * 2 * g.rbc + c;
    if (first >= g.nrb) return 0;
    const int cnt = (g.nrb - first + 1) / 2;
    return cnt < g.rbc ? cnt : g.rbc;
}
GPFO_HD int cg_kp_end(const ClusterGeom& g, int p) {
    const int e = (p + 1) * 2 * g.rbc;
    return e < g.nrb ? e : g.nrb;
}
// first stored local row block for diagonal offset kd = kp - 2*rbc*p: block (q, kp) is kept iff 2q + c >= kd
GPFO_HD int cg_lo(int kd, int c) {
    const int t = kd - c + 1;
    return t > 0 ? (t >> 1) : 0;
}

#endif  // GPFO_BUILD_EXPERIMENTS

// ------------------------------------------------------------------------------------------------
// Device-side description of one GP and of the acquisition program
// ------------------------------------------------------------------------------------------------
struct GpDev {
    int M, Mp, D, R;
    int kind;                   // GPFO_KERN_*
    int slot0;
    double variance;            // kernel variance (Kdiag)
    const double* Xs;           // Mp x D, training inputs divided by lengthscales (rows >= M are zero)
    const double* xn;           // Mp, squared norms of the rows of Xs
    const double* V;            // Mp x R, L^-1 Y (rows >= M are zero)
    const double* XsT;          // D x ldT, Xs transposed (thin.cu: lane = training point, coalesced)
    int ldT;
    const double* packed;       // block stream of L^-1
    PackGeom geom;
#ifdef GPFO_BUILD_EXPERIMENTS
    const double* packed_c[2];  // cluster layout: one block stream of L^-1 per CTA rank
    ClusterGeom cgeom;
#endif
    const double* alpha;        // Mp x R, K^-1 Y = L^-T V            (gradient pass only)
    const double* packed_kinv;  // dense block stream of K^-1         (gradient pass only)
    PackGeom geom_kinv;
    double in_scale[GPFO_MAX_DIM], in_shift[GPFO_MAX_DIM], inv_ell[GPFO_MAX_DIM];   // per input dimension
    double ell[GPFO_MAX_DIM];
    double out_scale[GPFO_MAX_SLOTS], out_shift[GPFO_MAX_SLOTS];                   // per output column
};

struct ProgramDev {
    int nops;
    int nslots;                 // moment slots referenced
    int has_hvpoi;
    int32_t ops[GPFO_MAX_OPS * 4];
    double params[GPFO_MAX_PARAMS];
};

struct EvalIO {
    const double* X;            // N x D candidates (device)
    int64_t N;
    long long mstride;          // slot stride of the [slot][.] arrays below (= candidates of the whole call)
    long long idx0;             // index of this launch's row 0 within the whole call (added to the argmax index)
    double* mom_mean;           // [slot][N] scratch (may be null when not needed)
    double* mom_var;
    double* values;             // N (may be null)
    double* part_val;           // per-CTA argmax partials
    long long* part_idx;
    int write_moments;          // this launch stores its moments to scratch
    int run_program;            // this launch evaluates the program in its epilogue
    // gradient pass
    double* cmu;                // [slot][N] d acq / d mean_slot   (written by the program-gradient kernel)
    double* cva;                // [slot][N] d acq / d var_slot
    double* grads;              // N x D  d acq / d Xcand
    int grad_accumulate;        // 0: this GP's launch overwrites grads, 1: adds to it
    // optional K(X*,X) scratch (variant 11): [grid][groups][GROUP_SLOTS] doubles, thread-private re-reads
    double* kx_scratch;
    long long kx_scratch_stride;   // doubles per CTA
    long long* debug;              // profiling builds only (ABL bit 64): per-warp phase cycle counters
    // on-device candidate generation: row r, column d = gen_b[d] + gen_A[d] * U(seed, (gen_first + r) * D + d)
    const double* gen_A;           // device, D values (null: candidates are read from X)
    const double* gen_b;
    unsigned long long gen_seed;
    long long gen_first;           // global index of this call's row 0 (lets every rank generate its own shard)
    // row-split launches (few candidates): block stream + layout of this launch, CTAs per tile, per-(tile, split) partial sums
    const double* packed_o;
    PackGeom geom_o;
    int nsplit;
    double* partials;              // forward: [tile][split][T][1+RMAX]; gradient pass: [tile][split][T][1+D]
    // stored-A gradient pass (large batches): the forward launch parks A = L^-1 Kx of every tile, reversed row order, in
    // B-fragment group layout; the backward launch contracts it with the reversed transposed stream: W = L^-T A
    double* a_scratch;             // [tile][groups][GROUP_SLOTS]
    long long a_stride;            // doubles per tile
};

// Counter-based uniform generator shared by the device prologue and the host regeneration (gpfo_random_rows):
// two rounds of the splitmix64 finaliser over (seed, element counter); the top 53 bits become a double in [0, 1).
GPFO_HD unsigned long long gpfo_mix64(unsigned long long x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
GPFO_HD double gpfo_uniform(unsigned long long seed, unsigned long long counter) {
    const unsigned long long bits = gpfo_mix64(seed ^ gpfo_mix64(counter + 0x9E3779B97F4A7C15ull));
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}

#define GPFO_STABILITY 1e-6     // gpflow settings.numerics.jitter_level (ei.py:24, poi.py:23, pof.py:23, hvpoi.py:24)

// ------------------------------------------------------------------------------------------------
// TF1 Normal.cdf / Normal.prob forms
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double gpfo_ndtr(double t) {
    const double half_sqrt2 = 0.70710678118654757;              // 0.5 * sqrt(2) as float64
    double w = t * half_sqrt2;
    double z = fabs(w);
    double y;
    if (z < half_sqrt2) y = 1.0 + erf(w);
    else if (w > 0.0)   y = 2.0 - erfc(z);
    else                y = erfc(z);
    return 0.5 * y;
}
__device__ __forceinline__ double gpfo_std_pdf(double z) {
    const double half_log_2pi = 0.91893853320467267;
    return exp(-0.5 * (z * z) - half_log_2pi);
}
// TF1 log_ndtr (Normal.log_cdf, float64 segments): -ndtr(-x) above 8, log(ndtr(x)) in (-20, 8], the 3-term asymptotic series
// below; d = derivative of the selected branch (what tf.gradients returns through tf.where).
__device__ __forceinline__ double gpfo_log_ndtr(double x, double* d) {
    const double half_log_2pi = 0.91893853320467267;
    if (x > 8.0) { if (d) *d = gpfo_std_pdf(x); return -gpfo_ndtr(-x); }
    if (x > -20.0) {
        const double c = gpfo_ndtr(x);
        if (d) *d = gpfo_std_pdf(x) / c;
        return log(c);
    }
    const double x2 = x * x, x4 = x2 * x2, x6 = x4 * x2;
    const double ser = 1.0 + 3.0 / x4 - (1.0 / x2 + 15.0 / x6);
    if (d) { const double x3 = x2 * x; *d = (-x - 1.0 / x) + (2.0 / x3 - 12.0 / (x3 * x2) + 90.0 / (x3 * x3 * x)) / ser; }
    return (-0.5 * x2 - log(-x) - half_log_2pi) + log(ser);
}
// Normal(loc, scale).prob(x) = exp(-0.5 z^2 - (0.5 log 2pi + log scale))
__device__ __forceinline__ double gpfo_normal_prob(double x, double loc, double scale) {
    const double half_log_2pi = 0.91893853320467267;
    double z = (x - loc) / scale;
    // explicit roundings (no FMA contraction): the same bits in every kernel this is inlined into, and TF's op order
    return exp(__dsub_rn(__dmul_rn(-0.5, __dmul_rn(z, z)), __dadd_rn(half_log_2pi, log(scale))));
}

// Evaluates the postfix acquisition program for one candidate.
//   mu, va   moment slots (un-scaled predictive mean / variance, NOT clamped)
//   hv       value of the HVPoI leaf for this candidate (only read when the program has one)
// With GRAD: also returns d value / d mu[s] and d value / d va[s] for every slot (forward mode over the
// tiny program; products use the product rule = TF's exclusive-product gradient).
template <bool GRAD>
__device__ __forceinline__ double gpfo_run_program(const ProgramDev& pr, const double* mu, const double* va,
                                                   double hv, const double* hv_dmu, const double* hv_dva,
                                                   double* dmu, double* dva) {
    constexpr int STK = 8;
    double st[STK];
    double gm[GRAD ? STK : 1][GRAD ? GPFO_MAX_SLOTS : 1];
    double gv[GRAD ? STK : 1][GRAD ? GPFO_MAX_SLOTS : 1];
    int sp = 0;
    const int ns = pr.nslots;
    for (int i = 0; i < pr.nops; ++i) {
        const int op = pr.ops[4 * i], a = pr.ops[4 * i + 1], b = pr.ops[4 * i + 2], c = pr.ops[4 * i + 3];
        if (op <= GPFO_OP_LCB || op == GPFO_OP_MEAN || op == GPFO_OP_VAR) {
            const double m = mu[a], v0 = va[a];
            double val, d_m = 0.0, d_v = 0.0;
            if (op == GPFO_OP_MEAN) { val = m; d_m = 1.0; }
            else if (op == GPFO_OP_VAR) { val = v0; d_v = 1.0; }
            else if (op == GPFO_OP_LCB) {
                const double sig = pr.params[c];
                const double v = fmax(v0, 0.0);
                const double s = sqrt(v);
                val = __dsub_rn(m, __dmul_rn(sig, s));
                if (GRAD) { d_m = 1.0; d_v = (v0 >= 0.0) ? -sig / (2.0 * s) : 0.0; }
            } else {
                const double p = pr.params[c];
                const double v = fmax(v0, GPFO_STABILITY);
                const double s = sqrt(v);
                const double z = (p - m) / s;
                const double Phi = gpfo_ndtr(z);
                const double ds_dv = (v0 >= GPFO_STABILITY) ? 1.0 / (2.0 * s) : 0.0;
                if (op == GPFO_OP_EI) {
                    val = __dadd_rn(__dmul_rn(p - m, Phi), __dmul_rn(v, gpfo_normal_prob(p, m, s)));
                    if (GRAD) { d_m = -Phi; d_v = gpfo_std_pdf(z) * ds_dv; }
                } else {
                    val = Phi;
                    if (GRAD) { const double phi = gpfo_std_pdf(z); d_m = -phi / s; d_v = -(z * phi / s) * ds_dv; }
                }
            }
            st[sp] = val;
            if (GRAD) {
                for (int s2 = 0; s2 < ns; ++s2) { gm[sp][s2] = 0.0; gv[sp][s2] = 0.0; }
                gm[sp][a] = d_m; gv[sp][a] = d_v;
            }
            ++sp;
        } else if (op == GPFO_OP_MES) {                         // mes.py:96-102; no variance clamp in the reference
            const double m = mu[a], v0 = va[a];
            const double s = sqrt(v0);
            double acc = 0.0, d_m = 0.0, d_v = 0.0;
            for (int k = 0; k < b; ++k) {
                const double g = (m - pr.params[c + k]) / s;
                const double pdf = gpfo_std_pdf(g), cdf = gpfo_ndtr(g);
                double dl = 0.0;
                const double lc = gpfo_log_ndtr(g, GRAD ? &dl : nullptr);
                acc += __dsub_rn(__dmul_rn(g, pdf) / __dmul_rn(2.0, cdf), lc);
                if (GRAD) {
                    const double dh = ((pdf - g * g * pdf) / (2.0 * cdf) - g * pdf * pdf / (2.0 * cdf * cdf)) - dl;
                    d_m += dh;
                    d_v += dh * (-g / (2.0 * v0));
                }
            }
            st[sp] = acc / (double)b;
            if (GRAD) {
                for (int s2 = 0; s2 < ns; ++s2) { gm[sp][s2] = 0.0; gv[sp][s2] = 0.0; }
                gm[sp][a] = d_m / s / (double)b; gv[sp][a] = d_v / (double)b;
            }
            ++sp;
        } else if (op == GPFO_OP_HVPOI) {
            st[sp] = hv;
            if (GRAD) {
                for (int s2 = 0; s2 < ns; ++s2) { gm[sp][s2] = 0.0; gv[sp][s2] = 0.0; }
                for (int j = 0; j < b; ++j) { gm[sp][a + j] = hv_dmu[j]; gv[sp][a + j] = hv_dva[j]; }
            }
            ++sp;
        } else if (op == GPFO_OP_SUM) {
            const int base = sp - a;
            double acc = st[base];
            for (int k = 1; k < a; ++k) acc += st[base + k];
            if (GRAD) {
                for (int k = 1; k < a; ++k)
                    for (int s2 = 0; s2 < ns; ++s2) { gm[base][s2] += gm[base + k][s2]; gv[base][s2] += gv[base + k][s2]; }
            }
            st[base] = acc; sp = base + 1;
        } else if (op == GPFO_OP_PROD) {
            const int base = sp - a;
            double acc = st[base];
            for (int k = 1; k < a; ++k) acc *= st[base + k];
            if (GRAD) {
                for (int s2 = 0; s2 < ns; ++s2) {
                    double tm = 0.0, tv = 0.0;
                    for (int k = 0; k < a; ++k) {
                        double others = 1.0;
                        for (int q = 0; q < a; ++q) if (q != k) others *= st[base + q];
                        tm += others * gm[base + k][s2];
                        tv += others * gv[base + k][s2];
                    }
                    gm[base][s2] = tm; gv[base][s2] = tv;       // operand k == 0 is overwritten last
                }
            }
            st[base] = acc; sp = base + 1;
        } else if (op == GPFO_OP_SCALE) {
            const double f = pr.params[c];
            st[sp - 1] *= f;
            if (GRAD) for (int s2 = 0; s2 < ns; ++s2) { gm[sp - 1][s2] *= f; gv[sp - 1][s2] *= f; }
        }
    }
    if (GRAD) for (int s2 = 0; s2 < ns; ++s2) { dmu[s2] = gm[0][s2]; dva[s2] = gv[0][s2]; }
    return st[0];
}

// (value, index) ordering used at every reduction level: larger value wins; ties -> lower index; NaN never wins.
__device__ __forceinline__ bool gpfo_better(double v, long long i, double bv, long long bi) {
    if (i < 0 || v != v) return false;                 // an empty or NaN challenger never wins ...
    if (bi < 0 || bv != bv) return true;               // ... and an empty or NaN incumbent always yields to a number
    return (v > bv) || (v == bv && i < bi);
}

// ------------------------------------------------------------------------------------------------
// host-side launchers implemented in the other translation units
// ------------------------------------------------------------------------------------------------
// factor.cu  (Mf = M rounded up to 64; all matrices have leading dimension Mf)
int gpfo_factor_device(cudaStream_t st, int M, int Mf, int D, int R, int kind, double variance, double noise,
                       const double* dXs, const double* dxn, double* dK /*Mf x Mf work, becomes L*/,
                       double* dLinv /*Mf x Mf*/, const double* dY /*Mf x R*/, double* dV /*Mf x R*/,
                       int* dflag, std::string* err);
void gpfo_pack_linv(cudaStream_t st, const double* dLinv, int ld, PackGeom g, double* dpacked);
#ifdef GPFO_BUILD_EXPERIMENTS
void gpfo_cluster_geom(int nrb, int rbc, ClusterGeom* g, long long total[2]);
void gpfo_pack_linv_cluster(cudaStream_t st, const double* dLinv, int ld, const ClusterGeom& g, double* dpacked0,
                            double* dpacked1);
// contract3.cu
cudaError_t gpfo_contract3_launch(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io,
                                  int nclusters, int D);
int gpfo_contract3_rbc();
int gpfo_contract3_clusters(int64_t N, int num_sms);
#endif
// contract5.cu (resident-panel kernel)
int gpfo_contract5_max_dim(int variant);
cudaError_t gpfo_contract5_astore_launch(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io,
                                         int num_sms, int D, int native, int variant);
cudaError_t gpfo_contract5_bwd_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int num_sms, int D);
int gpfo_contract5_bwd_ok(int D);
long long gpfo_contract5_astore_stride(int nrb);
// Kinv = Linv^T Linv (Mf x Mf, full symmetric) and alpha = Linv^T V, for the gradient pass
void gpfo_kinv_alpha(cudaStream_t st, const double* dLinv, int Mf, const double* dV, int R, double* dKinv, double* dalpha);
void gpfo_prepare_inputs(cudaStream_t st, const double* dX, int M, int Mp, int D, const double* d_ell,
                         double* dXs, double* dxn);

// contract.cu
int gpfo_contract_rbc(int variant);
int gpfo_contract_tile(int variant);
int gpfo_contract_grid(int variant, int64_t N, int num_sms);
cudaError_t gpfo_contract_launch(cudaStream_t st, int variant, int kind, const GpDev* d_gp, const ProgramDev* d_prog,
                                 EvalIO io, int grid, int D);
void gpfo_argmax_final(cudaStream_t st, const double* part_val, const long long* part_idx, int nparts,
                       double* out_val, long long* out_idx);
double gpfo_dmma_peak(cudaStream_t st, int num_sms);
int gpfo_fp64_pipes(cudaStream_t st, int num_sms, int mode, double* out);

// hvpoi.cu
struct CellsDev {
    int n_ext, R;
    int64_t C;
    const double* pf_ext;       // n_ext x R
    const int32_t* lb;          // C x R
    const int32_t* ub;
    const int32_t* off;         // C x R x 2: Phi-table entries (row * R + objective) of the upper / lower bound   (hvpoi2_kernel)
    const double* bnd;          // C x R x 2: pf_ext values of the upper / lower bound
};
int gpfo_hvpoi_grid(int64_t N, int num_sms, int n_ext, int R, int with_grad);
size_t gpfo_hvpoi_smem_bytes(int n_ext, int R);
cudaError_t gpfo_hvpoi_launch(cudaStream_t st, const ProgramDev* d_prog, const CellsDev& cells, EvalIO io, int grid,
                              int slot_first, int with_grad);
size_t gpfo_hvpoi_smem_bytes_grad(int n_ext, int R);

// grad.cu
cudaError_t gpfo_program_grad_launch(cudaStream_t st, const ProgramDev* d_prog, EvalIO io, int grid);
int gpfo_program_grad_grid(int64_t N, int num_sms);
cudaError_t gpfo_contract_grad_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int grid, int D);
int gpfo_contract_grad_grid(int64_t N, int num_sms);
int gpfo_contract_grad_rbc();
// stored-A gradient pass for large batches (contract2.cu)
int gpfo_contract_bwd_rbc();
int gpfo_contract_bwd_tile();
long long gpfo_contract_astore_stride(int nrb);
size_t gpfo_contract_bwd_smem(int D);
cudaError_t gpfo_contract_astore_launch(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int grid, int D);
cudaError_t gpfo_contract_bwd_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int grid, int D);
// reversed transpose of the top-left n x n corner: dst[i][k] = src[n-1-k][n-1-i] (factor.cu)
void gpfo_reverse_transpose(cudaStream_t st, const double* src, int ld, int n, double* dst);
// row-split launches for few candidates (contract2.cu): contraction over a slice of the rows per CTA + a finish kernel
int gpfo_split_rbc();
int gpfo_split_tile();
cudaError_t gpfo_contract_split_launch(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io, int D);
cudaError_t gpfo_contract_grad_split_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int D, int wide);
// latency path for a handful of candidates (contract2_aux.cu): forward and gradient contractions side by side, one finish kernel
struct ThinFinalArgs {
    const GpDev* gp[GPFO_MAX_GPS];
    const double* part_f[GPFO_MAX_GPS];     // forward partials of GP g:  [tile][ns_f][8][1 + RMAX]
    const double* part_g[GPFO_MAX_GPS];     // gradient partials of GP g: [tile][ns_g][8 (1 + R)][1 + D]
    int ns_f[GPFO_MAX_GPS], ns_g[GPFO_MAX_GPS];
    int R[GPFO_MAX_GPS], slot0[GPFO_MAX_GPS];
    int ngps, D, want_grad;
    double* best_out;                       // (value, index) record of the call when one block covers it; may be host-mapped
};
cudaError_t gpfo_thin_forward_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int D);
cudaError_t gpfo_thin_grad_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int D);
int gpfo_thin_final_grid(int64_t N);
// thin.cu: contraction kernels for one tile of <= 8 candidates spread over all SMs (16 rows of the block stream per CTA)
int gpfo_thin_nsplit(int64_t N, int num_sms, int nrb);
int64_t gpfo_thin_max_n();
int gpfo_thin_panel_cols(int64_t N);
void gpfo_transpose_inputs(cudaStream_t st, const double* dXs, int rows, int D, double* dXsT);
cudaError_t gpfo_thin_contract_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int D, int grad);
cudaError_t gpfo_thin_final_launch(cudaStream_t st, const ProgramDev* d_prog, EvalIO io, const ThinFinalArgs& a);
