/*
 * gpfo.h -- C-ABI of libgpfo.so: B200-native batched acquisition evaluation (FP64, sm_100a).
 *
 * This is the drop-in boundary for GPflowOpt's data-parallel hot path.  The reference has no FFI of its
 * own (pure Python over GPflow 0.5 / TF1); each entry point below replaces the reference interface cited
 * beside it (paths relative to the GPflowOpt 0.1.1 tree).  The binding a maintainer adds on the reference
 * side is a ctypes stub -- see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns an int status (GPFO_OK == 0); gpfo_last_error(ctx) gives the message
 *   - plain pointers and sizes only; row-major float64; int32 cell indices; int64 candidate counts
 *   - the caller owns every buffer it passes; the library owns all device state inside the ctx
 *   - one ctx = one CUDA device + one stream; NOT thread-safe; calls are synchronous on return
 *   - pointer arguments documented "host or device" are classified with cudaPointerGetAttributes
 *   - there is NO CPU fallback: without a CUDA device gpfo_ctx_create fails with GPFO_ERR_CUDA
 */
#ifndef GPFO_H_
#define GPFO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPFO_VERSION 200

/* status codes */
enum {
    GPFO_OK = 0,
    GPFO_ERR_NOT_SPD = 1,       /* Cholesky pivot <= 0: what surfaces as tf.errors.InvalidArgumentError in
                                   the reference (caught at acquisition/acquisition.py:117, bo.py:51) */
    GPFO_ERR_BAD_SHAPE = 2,     /* AssertionError sites of the reference (shape / domain asserts) */
    GPFO_ERR_CUDA = 3,
    GPFO_ERR_UNSUPPORTED = 4,   /* model/kernels/transforms outside the path (no CPU fallback by design) */
    GPFO_ERR_BAD_PROGRAM = 5,
    GPFO_ERR_STATE = 6          /* e.g. eval before gp/program were set */
};

/* covariance functions of GPflow 0.5.0 kernels.py used by the reference's models
   (testing/utility.py:47,55,61-62; doc notebooks) */
enum { GPFO_KERN_RBF = 0, GPFO_KERN_MATERN52 = 1, GPFO_KERN_MATERN32 = 2 };

/* acquisition program: postfix (RPN) op stream.  Each op is 4 int32 words {opcode, a, b, c} */
enum {
    GPFO_OP_EI = 1,     /* a = moment slot, c = index into params[] of fmin      (acquisition/ei.py:70-79)  */
    GPFO_OP_POI = 2,    /* a = moment slot, c = index of fmin                    (acquisition/poi.py:63-67) */
    GPFO_OP_POF = 3,    /* a = moment slot, c = index of threshold               (acquisition/pof.py:81-85) */
    GPFO_OP_LCB = 4,    /* a = moment slot, c = index of sigma                   (acquisition/lcb.py:54-57) */
    GPFO_OP_HVPOI = 5,  /* a = first moment slot, b = number of objectives R     (acquisition/hvpoi.py:98-140);
                           uses the tables given to gpfo_cells_set */
    GPFO_OP_SUM = 6,    /* a = number of operands popped                         (acquisition.py:360-362, 368-379) */
    GPFO_OP_PROD = 7,   /* a = number of operands popped                         (acquisition.py:360-362, 382-393) */
    GPFO_OP_SCALE = 8,  /* c = index of constant factor (MCMC mean 1/n)          (acquisition.py:446-448) */
    GPFO_OP_MEAN = 9,   /* a = moment slot: pushes the predictive mean (identity epilogue; predict_f) */
    GPFO_OP_VAR = 10,   /* a = moment slot: pushes the predictive variance */
    GPFO_OP_MES = 11    /* a = moment slot, b = number of y* samples, c = index of the first sample in params[]
                           (acquisition/mes.py:96-102: mean over samples of g pdf(g)/(2 cdf(g)) - log cdf(g)) */
};

#define GPFO_MAX_GPS 16
#define GPFO_MAX_SLOTS 16          /* total output columns over all GPs (a "moment slot" = one (mean,var) pair) */
#define GPFO_MAX_OPS 64
#define GPFO_MAX_PARAMS 64
#define GPFO_MAX_DIM 64

typedef struct gpfo_ctx gpfo_ctx;

/* device-time breakdown of the last calls, milliseconds (CUDA events on the ctx stream) */
typedef struct gpfo_timings {
    double factor_ms;      /* last gpfo_gp_set: kernel matrix + Cholesky + inverse + pack */
    double h2d_ms;         /* last gpfo_eval: candidate upload (0 when candidates were device-resident) */
    double eval_ms;        /* last gpfo_eval: all kernels incl. argmax */
    double d2h_ms;         /* last gpfo_eval: result download */
    double hot_kernel_ms;  /* last gpfo_eval: sum over GPs of the DMMA contraction kernel alone */
    int64_t launches;      /* kernels launched by the last gpfo_eval */
    int64_t variant;       /* forward contraction kernel the last gpfo_eval selected (table in csrc/contract.cu) */
} gpfo_timings;

/* ---- lifetime -------------------------------------------------------------------------------------- */
int gpfo_version(void);
/* Creates a context on CUDA device `device`.  Replaces: AutoFlow's per-object tf.Graph + tf.Session
   cache (acquisition/acquisition.py:247-267; GPflow param.AutoFlow). */
int gpfo_ctx_create(int device, gpfo_ctx** out);
int gpfo_ctx_destroy(gpfo_ctx* ctx);
const char* gpfo_last_error(const gpfo_ctx* ctx);   /* valid until the next call on ctx; never NULL */
const char* gpfo_status_string(int status);

/* ---- model state (once per BO iteration) ------------------------------------------------------------ */
/* Uploads one GPR model as wrapped by DataScaler and factorises it on the device:
 *   K = k(X,X) + noise_var I ; L = chol(K) ; Linv = L^-1 ; V = Linv Y     (all FP64)
 * Replaces: the part of GPR.build_predict that does not depend on Xcand (GPflow 0.5.0 gpr.py, called
 * from scaling.py:189), which the reference re-executes inside every session.run, and the DataScaler
 * transforms (scaling.py:165-190, transforms.py:102-140, domain.py:89-93).
 *   X            M x D, ALREADY input-transformed (what DataScaler stores in wrapped.X), host
 *   Y            M x R, ALREADY output-transformed (wrapped.Y), host
 *   lengthscales D values if ard != 0, else 1 value
 *   in_scale/in_shift    D values: x' = x * in_scale + in_shift   (diagonal LinearTransform A, b)
 *   out_scale/out_shift  R values: forward y' = y * out_scale + out_shift; predictions are mapped back
 *                        mean = (mean' - out_shift) / out_scale, var = var' / out_scale^2
 *   slot0        index of the first moment slot this GP writes (slots slot0 .. slot0+R-1)
 * Errors: GPFO_ERR_NOT_SPD, GPFO_ERR_BAD_SHAPE, GPFO_ERR_CUDA. */
int gpfo_gp_set(gpfo_ctx* ctx, int gp_id, int64_t M, int D, int R, const double* X, const double* Y,
                int kernel_kind, int ard, const double* lengthscales, double kern_variance,
                double noise_variance, const double* in_scale, const double* in_shift,
                const double* out_scale, const double* out_shift, int slot0);
/* Drops all GPs with id >= n_gps (used when the acquisition tree shrinks). */
int gpfo_gp_count_set(gpfo_ctx* ctx, int n_gps);

/* Copies back device-side factor data for tests (any pointer may be NULL): L (M x M, lower, row-major),
   Linv (M x M), V (M x R). */
int gpfo_gp_get_factor(gpfo_ctx* ctx, int gp_id, double* L, double* Linv, double* V);

/* Acquisition tree as a postfix program.  Replaces: the TF graph assembled by build_acquisition
 * (acquisition/acquisition.py:124-125, 360-362; ei.py:70-79; poi.py:63-67; pof.py:81-85; lcb.py:54-57;
 * hvpoi.py:98-140).  ops = nops x 4 int32, params = nparams doubles. */
int gpfo_program_set(gpfo_ctx* ctx, const int32_t* ops, int nops, const double* params, int nparams);

/* Pareto cell tables for GPFO_OP_HVPOI.  Replaces: pf_ext / bounds.lb / bounds.ub feeds of
 * hvpoi.py:104-137 (tables produced host-side by pareto.py:146-255).
 *   pf_ext  (P+2) x R: row 0 = -inf, rows 1..P = front sorted on objective 0, last row = reference point
 *   lb, ub  C x R int32 row indices into pf_ext */
int gpfo_cells_set(gpfo_ctx* ctx, const double* pf_ext, int n_ext_rows, int R, const int32_t* lb,
                   const int32_t* ub, int64_t C);

/* ---- the hot path ------------------------------------------------------------------------------------ */
/* Scores N candidates with the current program.
 * Replaces: Acquisition.evaluate / evaluate_with_gradients (acquisition/acquisition.py:247-267) and, via
 * best_value/best_index, the np.argmin of MCOptimizer._optimize (optim.py:155-164; first occurrence wins;
 * the library maximises the acquisition, i.e. minimises bo.py:244-245's inverse_acquisition).
 *   Xcand        N x D float64, host or device, ORIGINAL (untransformed) coordinates
 *   out_values   N doubles or NULL (host or device)
 *   out_grads    N x D doubles or NULL (host or device); d acq / d Xcand
 *   best_value   host pointer or NULL: max over candidates of the acquisition value
 *   best_index   host pointer or NULL: lowest index attaining it (-1 when N == 0)
 * N == 0 is legal and touches no output buffer except best_index. */
int gpfo_eval(gpfo_ctx* ctx, const double* Xcand, int64_t N, int D, double* out_values, double* out_grads,
              double* best_value, int64_t* best_index);

/* Scores N candidates GENERATED ON THE DEVICE instead of read from memory (no N x D host array, no H2D copy).
 * Replaces: RandomDesign.generate() + evaluate of MCOptimizer._optimize (design.py:55-70, 83-94; optim.py:152-164).
 * Row r (global index first_row + r), column d is  lower[d] + (upper[d] - lower[d]) * U  with U in [0,1) from a
 * counter-based generator keyed by (seed, (first_row + r) * D + d): any shard of rows can be produced independently, and
 * gpfo_random_rows reproduces the same rows on the host bit for bit (used to return the winning point).
 * The stream differs from np.random.rand by construction, hence opt-in (MCOptimizer(..., device_rng=True)). */
int gpfo_eval_random(gpfo_ctx* ctx, uint64_t seed, int64_t first_row, int64_t N, int D, const double* lower,
                     const double* upper, double* out_values, double* best_value, int64_t* best_index);
/* Host regeneration of rows [first_row, first_row + n) of the stream above; needs no GPU and no ctx. */
int gpfo_random_rows(uint64_t seed, int64_t first_row, int64_t n, int D, const double* lower, const double* upper,
                     double* out);

/* Predictive mean and variance of one GP at N points (original coordinates), un-scaled outputs.
 * Replaces: DataScaler.predict_f (scaling.py:192-197) as used by _setup (ei.py:67, poi.py:60,
 * hvpoi.py:92).  mean, var: N x R, host or device, either may be NULL. */
int gpfo_predict(gpfo_ctx* ctx, int gp_id, const double* X, int64_t N, int D, double* mean, double* var);

int gpfo_timings_get(const gpfo_ctx* ctx, gpfo_timings* out);

/* Tuning knob for the contraction kernel.  variant 0 = automatic: calls of one to four candidates take the latency path
 * (16-row kernels, csrc/thin.cu), up to 4 * SMs candidates the row-split launches, beyond that the kernel with the least
 * waves x cost per wave -- the resident-panel kernels (19 / 20 / 21: 8 / 16 / 32 candidates per tile, K(X*,X) kept in shared
 * memory) or round 1's tilings (6 / 5 / 11 / 12: 8 / 16 / 32 / 64 candidates, the last two with a K(X*,X) scratch in global
 * memory); large gradient batches take the stored-A pass.  Explicit values pin one kernel (4, 5, 6, 11, 12, 19, 20, 21 in
 * every build; 10, 13, 14, 22-39, 41-43, 121-127 only with GPFO_BUILD_EXPERIMENTS=1: rejected kernels, ablations, phase
 * counters); see the table in csrc/contract.cu and DESIGN.md 4.0 / 4.1 / 8.
 * Environment switches: GPFO_STRICT_ONCHIP=1 (read once: never park K(X*,X) in global memory); GPFO_NO_THIN=1 /
 * GPFO_THIN_MAX_N (latency path off / its candidate limit); GPFO_NO_GRAPHS=1 (no CUDA-graph replay of small calls); per call
 * GPFO_GRAD_DENSE=1 (gradient pass over the dense K^-1 stream for every batch size), GPFO_ASTORE_MIN_N (stored-A threshold),
 * GPFO_ASTORE_BUDGET_MB (bound of the stored-A scratch). */
int gpfo_set_variant(gpfo_ctx* ctx, int variant);

/* The forward kernel gpfo_eval would select for a batch of N candidates with the models registered now (what variant 0 =
 * automatic resolves to).  A caller that scores one candidate set in several calls of different sizes (the pipelined multi-GPU
 * route of MCOptimizer) pins this variant for all of them: per-candidate results are then bit-identical whatever the call a
 * row lands in, so the first-occurrence argmax of optim.py:158 holds across calls. */
int gpfo_select_variant(gpfo_ctx* ctx, int64_t N, int* variant);

/* Measures this device's FP64 DMMA issue peak with a register-only mma.sync m8n8k4 loop; TFLOP/s. */
int gpfo_measure_dmma_peak(gpfo_ctx* ctx, double* tflops);

/* Host-side setup of HVPoI: cell decomposition of the non-dominated region for R >= 2 objectives.
 * Replaces: Pareto.divide_conquer_nd (pareto.py:173-236), a pure-Python stack loop in the reference.
 *   front   P x R, the Pareto set sorted ascending on objective 0 (pareto.py:131-144)
 *   order   P x R int32, np.argsort(front, axis=0) (ranks are taken from the caller so that ties break as in NumPy)
 *   lb, ub  caller buffers of `capacity` x R int32 (row indices into pf_ext); *count receives the number of cells.
 * Returns GPFO_ERR_BAD_SHAPE with *count set when capacity is too small (call again with a larger buffer). */
int gpfo_pareto_cells(const double* front, const int32_t* order, int P, int R, double threshold, int32_t* lb,
                      int32_t* ub, int64_t capacity, int64_t* count);

/* Host-side setup of HVPoI / BO result assembly: dominance counts of the non-dominated sort.
 * Replaces: pareto.non_dominated_sort (pareto.py:78-90; an n x n x R temporary in the reference).
 *   Y          n x R objective rows;   dominance[i] receives the number of rows dominating row i (front: == 0). */
int gpfo_non_dominated_sort(const double* Y, int64_t n, int R, int64_t* dominance);

/* Host-side candidate source of the multi-GPU Monte-Carlo route: n x D doubles of np.random.rand (MT19937, 53-bit) continued from
 * the caller's RandomState (key[624], *pos as np.random.get_state() returns them; both updated), mapped x * scale[d] + shift[d]
 * (either may be NULL) in the same pass.  Replaces: RandomDesign.generate (design.py:55-70, 83-94), bit for bit. */
int gpfo_mt19937_design_rows(uint32_t* key, int32_t* pos, int64_t n, int D, const double* scale, const double* shift,
                             double* out);

/* Host evaluation of the branch-free Normal CDF the HVPoI Phi table uses on the device (csrc/ndtr_fast.cuh; replaces
   tf.contrib.distributions.Normal.cdf at hvpoi.py:112-113).  Same source compiled for the host: lets the CPU tests pin its
   accuracy against scipy.special.ndtr without a GPU.  Needs no ctx. */
int gpfo_ndtr_fast_host(const double* z, int64_t n, double* out);

/* FP64 pipe microbenchmarks (register-only loops, 16 warps/SM): mode 0 = DMMA only, 1 = DFMA only,
   2 = half of the warps of every scheduler DMMA and half DFMA concurrently.  TFLOP/s of each instruction class. */
int gpfo_measure_fp64_pipes(gpfo_ctx* ctx, int mode, double* dmma_tflops, double* dfma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* GPFO_H_ */
