// C-ABI of libgpfo.so (see include/gpfo.h).  One ctx = one device + one stream; not thread-safe.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <nvtx3/nvToolsExt.h>
#include "gpfo_internal.cuh"
#include "ndtr_fast.cuh"

// NVTX ranges around the phases of a call (upload / factor / eval / argmax / download): visible in Nsight Systems timelines
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#define FB 64
#define GPFO_STAGE_BYTES (4 * 65536)          // pinned staging: four 64 KB areas (candidates, values, gradients, best record)
#define GPFO_STAGE_AREA 65536
#define RMAX_KERNEL 4

struct GpHost {
    bool valid = false;
    int M = 0, Mp = 0, Mf = 0, D = 0, R = 0, kind = 0, slot0 = 0;
    double *dX = nullptr, *dell = nullptr, *dXs = nullptr, *dxn = nullptr, *dY = nullptr, *dV = nullptr;
    double* dXsT = nullptr;                                 // Xs transposed (D x Mf): coalesced K(X*,X) generation of the latency path
    double *dL = nullptr, *dLinv = nullptr;
    double* dpacked_by[3] = {nullptr, nullptr, nullptr};    // forward block streams: [0] 256-, [1] 512-, [2] 1024-row chunks (or other)
    size_t cap_packed_by[3] = {0, 0, 0};
    bool packed_ready[3] = {false, false, false};
    int packed_slot_rbc[3] = {0, 0, 0};
    double *dalpha = nullptr, *dpacked_kinv = nullptr;      // gradient pass state, built lazily
    bool kinv_ready = false;
#ifdef GPFO_BUILD_EXPERIMENTS
    double* dpacked_c[2] = {nullptr, nullptr};              // cluster layout (contract3), built lazily
    size_t cap_packed_c[2] = {0, 0};
    bool cluster_ready = false;
#endif
    int packed_rbc = 0;
    // small-chunk copies of both block streams for the row-split launches (few candidates)
    double *dpacked_s = nullptr, *dpacked_kinv_s = nullptr;
    size_t cap_packed_s = 0, cap_kinv_s = 0;
    PackGeom geom_s, geom_kinv_s;
    // reversed transposed triangular stream of L^-1 for the stored-A backward pass (large gradient batches), built lazily
    double* dpacked_rt = nullptr; size_t cap_packed_rt = 0;
    PackGeom geom_rt;
    bool rt_ready = false;
    double* dpacked_rt5 = nullptr; size_t cap_packed_rt5 = 0;   // the same stream in 1024-row chunks (tall-chunk backward kernel)
    PackGeom geom_rt5;
    size_t cap_mf = 0, cap_d = 0, cap_r = 0, cap_alpha = 0, cap_kinv = 0;
    GpDev hdev;
    GpDev* ddev = nullptr;
};

struct gpfo_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    GpHost gps[GPFO_MAX_GPS];
    int ngps = 0;
    ProgramDev hprog;
    ProgramDev* dprog = nullptr;
    bool prog_set = false;
    // cells
    CellsDev cells;
    double* dpf = nullptr;
    int32_t *dlb = nullptr, *dub = nullptr, *dcoff = nullptr;
    double* dcbnd = nullptr;
    bool cells_set = false;
    // scratch
    double* dXc = nullptr; size_t cap_xc = 0;
    double* dvalues = nullptr; size_t cap_values = 0;
    double *dmom_mean = nullptr, *dmom_var = nullptr; size_t cap_mom = 0;
    double *dcmu = nullptr, *dcva = nullptr; size_t cap_coef = 0;
    double* dgrads = nullptr; size_t cap_grads = 0;
    double* dkx_scratch = nullptr; size_t cap_kx_scratch = 0;
    double* dpartials = nullptr; size_t cap_partials = 0;    // row-split partial sums
    double* dthin = nullptr; size_t cap_thin = 0;            // latency path: partial sums of all its concurrent launches
    cudaStream_t stream2 = nullptr;                          // latency path: the gradient contractions run next to the forward ones
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    double* da_scratch = nullptr; size_t cap_a_scratch = 0;  // stored-A gradient pass: A = L^-1 Kx of one slab of tiles
    double* dgen = nullptr;                                  // 2 * GPFO_MAX_DIM doubles: affine map of the device candidate generator
    bool gen_active = false; unsigned long long gen_seed = 0; long long gen_first = 0;
    double* dpart_val = nullptr; long long* dpart_idx = nullptr; int cap_parts = 0;
    double* dbest_val = nullptr; long long* dbest_idx = nullptr;   // one 16-byte record: value, index
    // pinned host staging for small transfers (candidates in, values / gradients / best record out): pageable copies of a
    // few KB cost ~10 us each through the driver's own staging, which is most of a latency-bound N = 1 call
    char* hstage = nullptr;
    int* dflag = nullptr;
    cudaEvent_t ev[8];
    cudaEvent_t gp_ev[GPFO_MAX_GPS][2];        // per-GP event pair around the contraction launch (read after the final sync)
    gpfo_timings tm;
    int variant = 0;
    // CUDA graphs of small calls (the N = 1 ... few hundred polish / setup calls of a BO iteration are launch bound: six kernel
    // launches, four small copies and a dozen event records per call).  The whole enqueue sequence of a call shape is captured
    // the second time the shape is seen and replayed afterwards; `epoch` changes with everything the sequence depends on.
    struct GraphEntry { unsigned long long key[4]; cudaGraphExec_t exec; int launches; int variant; };
    std::vector<GraphEntry> graphs;
    unsigned long long epoch = 0;
    bool capturing = false;
    long long graph_replays = 0;
};

static void drop_graphs(gpfo_ctx* ctx) {
    for (auto& g : ctx->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
}

static const char* k_status[] = {"ok", "matrix not positive definite", "bad shape", "CUDA error",
                                 "unsupported model or feature", "bad acquisition program", "invalid state"};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                              \
            return GPFO_ERR_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

static int fail(gpfo_ctx* ctx, int code, const std::string& msg) {
    ctx->err = msg;
    return code;
}

// Bumped whenever a device buffer is (re)allocated: cached CUDA graphs bake buffer addresses in, so they are keyed by it.
static unsigned long long g_alloc_epoch = 0;

template <typename T>
static cudaError_t ensure(T** p, size_t* cap, size_t need) {
    if (*cap >= need && *p) return cudaSuccess;
    ++g_alloc_epoch;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc((void**)p, need * sizeof(T));
    if (e == cudaSuccess) *cap = need;
    return e;
}

static bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

extern "C" {

int gpfo_version(void) { return GPFO_VERSION; }

const char* gpfo_status_string(int status) {
    if (status < 0 || status > GPFO_ERR_STATE) return "unknown status";
    return k_status[status];
}

const char* gpfo_last_error(const gpfo_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gpfo_ctx_create(int device, gpfo_ctx** out) {
    if (!out) return GPFO_ERR_BAD_SHAPE;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GPFO_ERR_CUDA; }
    if (device < 0 || device >= ndev) return GPFO_ERR_BAD_SHAPE;
    if (cudaSetDevice(device) != cudaSuccess) return GPFO_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GPFO_ERR_CUDA;
    if (prop.major != 10) {
        fprintf(stderr, "libgpfo: device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major,
                prop.minor);
        return GPFO_ERR_UNSUPPORTED;
    }
    gpfo_ctx* ctx = new gpfo_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    memset(&ctx->hprog, 0, sizeof(ctx->hprog));
    memset(&ctx->cells, 0, sizeof(ctx->cells));
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 8 && ok; ++i) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
    for (int i = 0; i < 2 * GPFO_MAX_GPS && ok; ++i) ok = cudaEventCreate(&ctx->gp_ev[i / 2][i % 2]) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&ctx->dprog, sizeof(ProgramDev)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&ctx->dbest_val, 16) == cudaSuccess;
    if (ok) ctx->dbest_idx = reinterpret_cast<long long*>(ctx->dbest_val + 1);
    ok = ok && cudaMallocHost((void**)&ctx->hstage, GPFO_STAGE_BYTES) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&ctx->dflag, sizeof(int)) == cudaSuccess;
    if (!ok) { delete ctx; return GPFO_ERR_CUDA; }
    *out = ctx;
    return GPFO_OK;
}

static void free_gp(GpHost& g) {
    cudaFree(g.dX); cudaFree(g.dell); cudaFree(g.dXs); cudaFree(g.dXsT); cudaFree(g.dxn); cudaFree(g.dY); cudaFree(g.dV);
    cudaFree(g.dL); cudaFree(g.dLinv); cudaFree(g.dpacked_by[0]); cudaFree(g.dpacked_by[1]); cudaFree(g.dpacked_by[2]); cudaFree(g.ddev);
    cudaFree(g.dalpha); cudaFree(g.dpacked_kinv);
#ifdef GPFO_BUILD_EXPERIMENTS
    cudaFree(g.dpacked_c[0]); cudaFree(g.dpacked_c[1]);
#endif
    cudaFree(g.dpacked_s); cudaFree(g.dpacked_kinv_s); cudaFree(g.dpacked_rt); cudaFree(g.dpacked_rt5);
    g = GpHost();
}

int gpfo_ctx_destroy(gpfo_ctx* ctx) {
    if (!ctx) return GPFO_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    drop_graphs(ctx);
    for (int i = 0; i < GPFO_MAX_GPS; ++i) free_gp(ctx->gps[i]);
    cudaFree(ctx->dprog); cudaFree(ctx->dpf); cudaFree(ctx->dlb); cudaFree(ctx->dub); cudaFree(ctx->dcoff); cudaFree(ctx->dcbnd);
    cudaFree(ctx->dXc); cudaFree(ctx->dvalues); cudaFree(ctx->dmom_mean); cudaFree(ctx->dmom_var);
    cudaFree(ctx->dcmu); cudaFree(ctx->dcva); cudaFree(ctx->dgrads); cudaFree(ctx->dkx_scratch); cudaFree(ctx->dgen);
    cudaFree(ctx->dpartials); cudaFree(ctx->da_scratch); cudaFree(ctx->dthin);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    cudaFree(ctx->dpart_val); cudaFree(ctx->dpart_idx); cudaFree(ctx->dbest_val); cudaFreeHost(ctx->hstage);
    cudaFree(ctx->dflag);
    for (int i = 0; i < 8; ++i) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 2 * GPFO_MAX_GPS; ++i) cudaEventDestroy(ctx->gp_ev[i / 2][i % 2]);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return GPFO_OK;
}

// Relative cost of one large-batch pass over Mp training points with a given forward kernel: block-steps the kernel issues
// (its last, ragged chunk is padded to whole quads of row blocks per warp in v5) over its measured DMMA efficiency on full
// chunks (profiles/r02_v5_sweep_3_pipelined_tiles.log: v20 69-82 %, v21 55-72 %, v12 51-76 %).
static double large_batch_cost(int variant, int Mp) {
    const int nrb = Mp / 8;
    // measured fractions of the DMMA peak (profiles/r02_v5_sweep_final.log)
    if (variant == 12) return (double)nrb * (nrb + 1) / 2 / (Mp >= 1536 ? 0.74 : (Mp >= 768 ? 0.65 : 0.51));
    const int rbc = (variant == 20 || variant == 19) ? 128 : 64;
    double steps = 0.0;
    for (int row0 = 0; row0 < nrb; row0 += rbc) {
        const int nr = nrb - row0 < rbc ? nrb - row0 : rbc;
        if (nr == rbc) steps += (double)nr * row0 + (double)nr * (nr + 1) / 2;
        else                                        // short last chunk: run as a full one whose missing row blocks are zero rows
            steps += (double)rbc * row0 + (double)rbc * nr - (double)nr * (nr - 1) / 2;
    }
    // (variant 19 = variant 20's tiling with 8 candidates: 0.8 of its efficiency, profiles/r02_midsize_batches.log)
    if (variant == 19) return steps / (0.8 * (Mp >= 1536 ? 0.815 : 0.725));
    return steps / (variant == 20 ? (Mp >= 1536 ? 0.815 : 0.725) : (Mp >= 1536 ? 0.75 : (Mp >= 768 ? 0.70 : 0.625)));
}

static int effective_variant(const gpfo_ctx* ctx, int64_t N, int Mp) {
    if (ctx->variant >= 1) return ctx->variant;
    // automatic.  Large batches (>= four waves of 64-candidate tiles): the resident-panel kernels (v5: variant 20 = 16
    // candidates x 1024-row chunks, 21 = 32 x 512; K(X*,X) never leaves the SM) unless the training-set size leaves a mostly
    // empty last chunk, where the 64-candidate tiles with the K(X*,X) scratch (variant 12) do less padded work;
    // GPFO_STRICT_ONCHIP=1 rules the scratch out altogether.  Below that wave quantisation decides: cost = waves of tiles x
    // measured time of one wave relative to the others (M = 2048: 0.29 / 0.50 / 0.77 (0.85 on chip) / 1.46 ms for
    // T = 8 / 16 / 32 / 64); ties go to the wider tile.
    static const bool strict = getenv("GPFO_STRICT_ONCHIP") && atoi(getenv("GPFO_STRICT_ONCHIP")) != 0;
    if (N >= (int64_t)64 * 4 * ctx->num_sms) {
        int best = strict ? 20 : 12;
        double best_cost = large_batch_cost(best, Mp);
        for (int v : {20, 21}) {
            const double c = large_batch_cost(v, Mp);
            if (c < best_cost) { best = v; best_cost = c; }
        }
        return best;
    }
    const int var[4] = {6, 5, strict ? 4 : 11, 12};
    const int tile[4] = {8, 16, 32, 64};
    const double wave_cost[4] = {0.29, 0.50, strict ? 0.85 : 0.77, 1.46};
    int best = 0;
    double best_cost = 0.0;
    for (int k = 0; k < (strict ? 3 : 4); ++k) {
        const int64_t tiles = (N + tile[k] - 1) / tile[k];
        const double cost = (double)((tiles + ctx->num_sms - 1) / ctx->num_sms) * wave_cost[k];
        if (k == 0 || cost <= best_cost) { best = k; best_cost = cost; }
    }
    int best_var = var[best];
    // ... and the resident-panel kernels, whose wave of 16 / 32 candidates costs what the large-batch model says relative to a
    // wave of the 64-candidate tiles (M = 2048: 0.33 ms per wave of 16 against 0.50 ms for the round-1 tiling of 16): from a few
    // waves of their tiles on they win the mid-size batches as well (MCOptimizer(domain, 10^4) at M = 2048: 2.3 -> 1.7 ms)
    const double c12 = large_batch_cost(12, Mp);
    for (int v : {19, 20, 21}) {
        const int T = v == 19 ? 8 : (v == 20 ? 16 : 32);
        const int64_t tiles = (N + T - 1) / T;
        const double cost = (double)((tiles + ctx->num_sms - 1) / ctx->num_sms) * 1.46 * (T * large_batch_cost(v, Mp)) / (64.0 * c12);
        if (cost < best_cost) { best_var = v; best_cost = cost; }
    }
    return best_var;
}

static PackGeom make_geom(int nrb, int rbc, int dense) {
    PackGeom geom;
    geom.nrb = nrb;
    geom.rbc = rbc;
    geom.nchunks = (nrb + rbc - 1) / rbc;
    geom.dense = dense;
    return geom;
}

// Row split: with few candidate tiles the rows of the block stream are spread over the SMs (nsplit CTAs per tile).
// Returns the CTAs per tile, 1 = no split.  Explicit kernel variants (gpfo_set_variant) keep their own launch shape.
static int split_count(const gpfo_ctx* ctx, int64_t N, int T, int nchunks) {
    if (ctx->variant != 0) return 1;
    const int64_t ntiles = (N + T - 1) / T;
    if (ntiles * 2 > ctx->num_sms || nchunks < 2) return 1;
    const int ns = (int)(ctx->num_sms / ntiles);
    return ns < nchunks ? ns : nchunks;
}

// Dense gradient pass (32-candidate tiles): CTAs per tile chosen so that whole waves are filled -- each CTA contracts 1/ns of
// the rows, so the cost is waves(tiles * ns) / ns, plus a few percent per split for the partial sums and the finish kernel.
static int grad_split_wide(const gpfo_ctx* ctx, int64_t N, int nchunks) {
    if (ctx->variant != 0) return 1;
    const char* off = getenv("GPFO_GRAD_NO_WAVE_SPLIT");
    const int64_t ntiles = (N + 31) / 32;
    if (off && atoi(off) != 0) return (ntiles * 2 > ctx->num_sms || nchunks < 2) ? 1 : (int)std::min<int64_t>(ctx->num_sms / ntiles, nchunks);
    int best = 1;
    double best_cost = (double)((ntiles + ctx->num_sms - 1) / ctx->num_sms);
    for (int ns = 2; ns <= nchunks && ns <= 16 && ntiles * ns <= (int64_t)4 * ctx->num_sms; ++ns) {
        const double cost = (double)((ntiles * ns + ctx->num_sms - 1) / ctx->num_sms) / ns * (1.0 + 0.03 * (ns - 1));
        if (cost < best_cost) { best = ns; best_cost = cost; }
    }
    return best;
}

static int ensure_partials(gpfo_ctx* ctx) {
    CK(ensure(&ctx->dpartials, &ctx->cap_partials, (size_t)4 * ctx->num_sms * 32 * (1 + GPFO_MAX_DIM)));
    return GPFO_OK;
}

// Block stream of L^-1 with `rbc` row blocks per chunk, built on first use and cached in its slot; does not change which
// layout the GP's device record points at.
static int ensure_pack(gpfo_ctx* ctx, GpHost& g, int rbc, const double** packed, PackGeom* geom_out) {
    const PackGeom geom = make_geom(g.Mp / 8, rbc, 0);
    const int slot = rbc == 32 ? 0 : (rbc == 64 ? 1 : 2);
    if (!g.packed_ready[slot] || g.packed_slot_rbc[slot] != rbc) {
        CK(ensure(&g.dpacked_by[slot], &g.cap_packed_by[slot], (size_t)pg_total_blocks(geom) * 64));
        gpfo_pack_linv(ctx->stream, g.dLinv, g.Mf, geom, g.dpacked_by[slot]);
        CK(cudaGetLastError());
        g.packed_ready[slot] = true;
        g.packed_slot_rbc[slot] = rbc;
    }
    *packed = g.dpacked_by[slot];
    *geom_out = geom;
    return GPFO_OK;
}

// Makes the block stream with `rbc` row blocks per chunk the GP's current one.  Both layouts in use (256- and 512-row chunks)
// stay cached, so alternating batch sizes only re-sends the small GpDev record.
static int repack(gpfo_ctx* ctx, GpHost& g, int rbc) {
    const PackGeom geom = make_geom(g.Mp / 8, rbc, 0);
    const int slot = rbc == 32 ? 0 : (rbc == 64 ? 1 : 2);
    if (!g.packed_ready[slot] || g.packed_slot_rbc[slot] != rbc) {
        const size_t need = (size_t)pg_total_blocks(geom) * 64;
        CK(ensure(&g.dpacked_by[slot], &g.cap_packed_by[slot], need));
        gpfo_pack_linv(ctx->stream, g.dLinv, g.Mf, geom, g.dpacked_by[slot]);
        CK(cudaGetLastError());
        g.packed_ready[slot] = true;
        g.packed_slot_rbc[slot] = rbc;
    }
    ++ctx->epoch;
    g.packed_rbc = rbc;
    g.hdev.geom = geom;
    g.hdev.packed = g.dpacked_by[slot];
    CK(cudaMemcpyAsync(g.ddev, &g.hdev, sizeof(GpDev), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GPFO_OK;
}

#ifdef GPFO_BUILD_EXPERIMENTS
// Cluster-layout block streams of L^-1 for the 2-CTA kernel (variant 10), built on first use.
static int ensure_cluster_pack(gpfo_ctx* ctx, GpHost& g) {
    if (g.cluster_ready) return GPFO_OK;
    ClusterGeom cgm;
    long long total[2];
    gpfo_cluster_geom(g.Mp / 8, gpfo_contract3_rbc(), &cgm, total);
    if (cgm.nchunks > GPFO_MAX_CCHUNKS) return fail(ctx, GPFO_ERR_UNSUPPORTED, "M too large for the cluster layout");
    for (int c = 0; c < 2; ++c) CK(ensure(&g.dpacked_c[c], &g.cap_packed_c[c], (size_t)(total[c] > 0 ? total[c] : 1) * 64));
    gpfo_pack_linv_cluster(ctx->stream, g.dLinv, g.Mf, cgm, g.dpacked_c[0], g.dpacked_c[1]);
    CK(cudaGetLastError());
    g.hdev.packed_c[0] = g.dpacked_c[0];
    g.hdev.packed_c[1] = g.dpacked_c[1];
    g.hdev.cgeom = cgm;
    CK(cudaMemcpyAsync(g.ddev, &g.hdev, sizeof(GpDev), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    g.cluster_ready = true;
    return GPFO_OK;
}
#endif  // GPFO_BUILD_EXPERIMENTS

// Gradient-pass state of one GP: alpha = L^-T V and the dense block stream of K^-1 = L^-T L^-1 (built on first use).
static int ensure_kinv(gpfo_ctx* ctx, GpHost& g) {
    if (g.kinv_ready) return GPFO_OK;
    cudaStream_t st = ctx->stream;
    PackGeom geom;
    geom.nrb = g.Mp / 8;
    geom.rbc = gpfo_contract_grad_rbc();
    geom.nchunks = (geom.nrb + geom.rbc - 1) / geom.rbc;
    geom.dense = 1;
    g.geom_kinv_s = make_geom(geom.nrb, gpfo_split_rbc(), 1);
    CK(ensure(&g.dalpha, &g.cap_alpha, (size_t)g.Mf * g.R));
    CK(ensure(&g.dpacked_kinv, &g.cap_kinv, (size_t)pg_total_blocks(geom) * 64));
    CK(ensure(&g.dpacked_kinv_s, &g.cap_kinv_s, (size_t)pg_total_blocks(g.geom_kinv_s) * 64));
    double* dKinv = nullptr;
    CK(cudaMalloc((void**)&dKinv, (size_t)g.Mf * g.Mf * sizeof(double)));
    gpfo_kinv_alpha(st, g.dLinv, g.Mf, g.dV, g.R, dKinv, g.dalpha);
    gpfo_pack_linv(st, dKinv, g.Mf, geom, g.dpacked_kinv);
    gpfo_pack_linv(st, dKinv, g.Mf, g.geom_kinv_s, g.dpacked_kinv_s);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(dKinv);
    if (e != cudaSuccess) return fail(ctx, GPFO_ERR_CUDA, std::string("ensure_kinv: ") + cudaGetErrorString(e));
    g.hdev.alpha = g.dalpha;
    g.hdev.packed_kinv = g.dpacked_kinv;
    g.hdev.geom_kinv = geom;
    CK(cudaMemcpy(g.ddev, &g.hdev, sizeof(GpDev), cudaMemcpyHostToDevice));
    g.kinv_ready = true;
    return GPFO_OK;
}

// Stored-A backward pass state: block stream of the reversed transpose of L^-1 (built on first use).
static int ensure_rt(gpfo_ctx* ctx, GpHost& g) {
    if (g.rt_ready) return GPFO_OK;
    cudaStream_t st = ctx->stream;
    g.geom_rt = make_geom(g.Mp / 8, gpfo_contract_bwd_rbc(), 0);
    g.geom_rt5 = make_geom(g.Mp / 8, gpfo_contract_rbc(20), 0);
    CK(ensure(&g.dpacked_rt, &g.cap_packed_rt, (size_t)pg_total_blocks(g.geom_rt) * 64));
    CK(ensure(&g.dpacked_rt5, &g.cap_packed_rt5, (size_t)pg_total_blocks(g.geom_rt5) * 64));
    double* tmp = nullptr;
    CK(cudaMalloc((void**)&tmp, (size_t)g.Mf * g.Mf * sizeof(double)));
    gpfo_reverse_transpose(st, g.dLinv, g.Mf, g.Mp, tmp);
    gpfo_pack_linv(st, tmp, g.Mf, g.geom_rt, g.dpacked_rt);
    gpfo_pack_linv(st, tmp, g.Mf, g.geom_rt5, g.dpacked_rt5);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(ctx, GPFO_ERR_CUDA, std::string("ensure_rt: ") + cudaGetErrorString(e));
    g.rt_ready = true;
    return GPFO_OK;
}

int gpfo_gp_set(gpfo_ctx* ctx, int gp_id, int64_t M64, int D, int R, const double* X, const double* Y,
                int kernel_kind, int ard, const double* lengthscales, double kern_variance, double noise_variance,
                const double* in_scale, const double* in_shift, const double* out_scale, const double* out_shift,
                int slot0) {
    if (!ctx) return GPFO_ERR_STATE;
    cudaSetDevice(ctx->device);
    if (gp_id < 0 || gp_id >= GPFO_MAX_GPS) return fail(ctx, GPFO_ERR_BAD_SHAPE, "gp_id out of range");
    if (M64 < 1 || M64 > 65536) return fail(ctx, GPFO_ERR_BAD_SHAPE, "M must be in [1, 65536]");
    if (D < 1 || D > GPFO_MAX_DIM) return fail(ctx, GPFO_ERR_BAD_SHAPE, "D must be in [1, 64]");
    if (R < 1 || R > RMAX_KERNEL) return fail(ctx, GPFO_ERR_UNSUPPORTED, "R (output columns per GP) must be in [1, 4]");
    if (slot0 < 0 || slot0 + R > GPFO_MAX_SLOTS) return fail(ctx, GPFO_ERR_BAD_SHAPE, "moment slot out of range");
    if (kernel_kind < GPFO_KERN_RBF || kernel_kind > GPFO_KERN_MATERN32)
        return fail(ctx, GPFO_ERR_UNSUPPORTED, "kernel must be RBF, Matern52 or Matern32");
    if (!X || !Y || !lengthscales) return fail(ctx, GPFO_ERR_BAD_SHAPE, "null input");
    if (!(kern_variance > 0.0) || !(noise_variance >= 0.0)) return fail(ctx, GPFO_ERR_BAD_SHAPE, "variances must be positive");
    const int M = (int)M64;
    ++ctx->epoch;
    GpHost& g = ctx->gps[gp_id];
    const int Mp = (M + 7) / 8 * 8, Mf = (M + FB - 1) / FB * FB;
    if ((size_t)Mf > g.cap_mf || (size_t)D > g.cap_d || (size_t)R > g.cap_r) {
        free_gp(g);
        size_t c = 0;
        CK(ensure(&g.dX, &c, (size_t)Mf * D)); c = 0;
        CK(ensure(&g.dell, &c, (size_t)D)); c = 0;
        CK(ensure(&g.dXs, &c, (size_t)Mf * D)); c = 0;
        CK(ensure(&g.dXsT, &c, (size_t)Mf * D)); c = 0;
        CK(ensure(&g.dxn, &c, (size_t)Mf)); c = 0;
        CK(ensure(&g.dY, &c, (size_t)Mf * R)); c = 0;
        CK(ensure(&g.dV, &c, (size_t)Mf * R)); c = 0;
        CK(ensure(&g.dL, &c, (size_t)Mf * Mf)); c = 0;
        CK(ensure(&g.dLinv, &c, (size_t)Mf * Mf));
        CK(cudaMalloc((void**)&g.ddev, sizeof(GpDev)));
        g.cap_mf = Mf; g.cap_d = D; g.cap_r = R;
    }
    g.valid = false;
    g.kinv_ready = false;
#ifdef GPFO_BUILD_EXPERIMENTS
    g.cluster_ready = false;
#endif
    g.rt_ready = false;
    g.packed_ready[0] = g.packed_ready[1] = g.packed_ready[2] = false;
    g.M = M; g.Mp = Mp; g.Mf = Mf; g.D = D; g.R = R; g.kind = kernel_kind; g.slot0 = slot0;
    memset(&g.hdev, 0, sizeof(GpDev));
    double ell[GPFO_MAX_DIM];
    for (int d = 0; d < D; ++d) {
        ell[d] = ard ? lengthscales[d] : lengthscales[0];
        if (!(ell[d] > 0.0)) return fail(ctx, GPFO_ERR_BAD_SHAPE, "lengthscales must be positive");
        g.hdev.ell[d] = ell[d];
        g.hdev.inv_ell[d] = 1.0 / ell[d];
        g.hdev.in_scale[d] = in_scale ? in_scale[d] : 1.0;
        g.hdev.in_shift[d] = in_shift ? in_shift[d] : 0.0;
    }
    for (int r = 0; r < R; ++r) {
        g.hdev.out_scale[r] = out_scale ? out_scale[r] : 1.0;
        g.hdev.out_shift[r] = out_shift ? out_shift[r] : 0.0;
        if (g.hdev.out_scale[r] == 0.0) return fail(ctx, GPFO_ERR_BAD_SHAPE, "out_scale must be non-zero");
    }
    cudaStream_t st = ctx->stream;
    NvtxRange nvtx_factor("gpfo:factor");
    CK(cudaEventRecord(ctx->ev[0], st));
    CK(cudaMemsetAsync(g.dX, 0, (size_t)Mf * D * sizeof(double), st));
    CK(cudaMemsetAsync(g.dY, 0, (size_t)Mf * R * sizeof(double), st));
    CK(cudaMemcpyAsync(g.dX, X, (size_t)M * D * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g.dY, Y, (size_t)M * R * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g.dell, ell, (size_t)D * sizeof(double), cudaMemcpyHostToDevice, st));
    gpfo_prepare_inputs(st, g.dX, M, Mf, D, g.dell, g.dXs, g.dxn);
    gpfo_transpose_inputs(st, g.dXs, Mf, D, g.dXsT);
    CK(cudaGetLastError());
    int rc = gpfo_factor_device(st, M, Mf, D, R, kernel_kind, kern_variance, noise_variance, g.dXs, g.dxn, g.dL,
                                g.dLinv, g.dY, g.dV, ctx->dflag, &ctx->err);
    if (rc != GPFO_OK) return rc;
    g.hdev.M = M; g.hdev.Mp = Mp; g.hdev.D = D; g.hdev.R = R; g.hdev.kind = kernel_kind; g.hdev.slot0 = slot0;
    g.hdev.variance = kern_variance;
    g.hdev.Xs = g.dXs; g.hdev.xn = g.dxn; g.hdev.V = g.dV;
    g.hdev.XsT = g.dXsT; g.hdev.ldT = Mf;
    g.geom_s = make_geom(Mp / 8, gpfo_split_rbc(), 0);
    CK(ensure(&g.dpacked_s, &g.cap_packed_s, (size_t)pg_total_blocks(g.geom_s) * 64));
    gpfo_pack_linv(st, g.dLinv, g.Mf, g.geom_s, g.dpacked_s);
    rc = repack(ctx, g, gpfo_contract_rbc(ctx->variant));
    if (rc != GPFO_OK) return rc;
    CK(cudaEventRecord(ctx->ev[1], st));
    CK(cudaEventSynchronize(ctx->ev[1]));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->tm.factor_ms = ms;
    g.valid = true;
    if (gp_id + 1 > ctx->ngps) ctx->ngps = gp_id + 1;
    return GPFO_OK;
}

int gpfo_gp_count_set(gpfo_ctx* ctx, int n_gps) {
    if (!ctx) return GPFO_ERR_STATE;
    if (n_gps < 0 || n_gps > GPFO_MAX_GPS) return fail(ctx, GPFO_ERR_BAD_SHAPE, "n_gps out of range");
    cudaSetDevice(ctx->device);
    ++ctx->epoch;
    for (int i = n_gps; i < GPFO_MAX_GPS; ++i)
        if (ctx->gps[i].valid || ctx->gps[i].dL) free_gp(ctx->gps[i]);
    ctx->ngps = n_gps;
    return GPFO_OK;
}

int gpfo_gp_get_factor(gpfo_ctx* ctx, int gp_id, double* L, double* Linv, double* V) {
    if (!ctx) return GPFO_ERR_STATE;
    if (gp_id < 0 || gp_id >= GPFO_MAX_GPS || !ctx->gps[gp_id].valid) return fail(ctx, GPFO_ERR_STATE, "gp not set");
    cudaSetDevice(ctx->device);
    GpHost& g = ctx->gps[gp_id];
    const size_t w = (size_t)g.M * sizeof(double);
    if (L) CK(cudaMemcpy2D(L, w, g.dL, (size_t)g.Mf * sizeof(double), w, g.M, cudaMemcpyDeviceToHost));
    if (Linv) CK(cudaMemcpy2D(Linv, w, g.dLinv, (size_t)g.Mf * sizeof(double), w, g.M, cudaMemcpyDeviceToHost));
    if (V) CK(cudaMemcpy(V, g.dV, (size_t)g.M * g.R * sizeof(double), cudaMemcpyDeviceToHost));
    return GPFO_OK;
}

int gpfo_program_set(gpfo_ctx* ctx, const int32_t* ops, int nops, const double* params, int nparams) {
    if (!ctx) return GPFO_ERR_STATE;
    cudaSetDevice(ctx->device);
    if (nops < 1 || nops > GPFO_MAX_OPS || nparams < 0 || nparams > GPFO_MAX_PARAMS || !ops)
        return fail(ctx, GPFO_ERR_BAD_PROGRAM, "program size out of range");
    ++ctx->epoch;
    ProgramDev p;
    memset(&p, 0, sizeof(p));
    int sp = 0, nslots = 0, hv = 0;
    for (int i = 0; i < nops; ++i) {
        const int op = ops[4 * i], a = ops[4 * i + 1], b = ops[4 * i + 2], c = ops[4 * i + 3];
        switch (op) {
            case GPFO_OP_EI: case GPFO_OP_POI: case GPFO_OP_POF: case GPFO_OP_LCB:
                if (c < 0 || c >= nparams) return fail(ctx, GPFO_ERR_BAD_PROGRAM, "leaf parameter index out of range");
                /* fallthrough */
            case GPFO_OP_MEAN: case GPFO_OP_VAR:
                if (a < 0 || a >= GPFO_MAX_SLOTS) return fail(ctx, GPFO_ERR_BAD_PROGRAM, "moment slot out of range");
                if (a + 1 > nslots) nslots = a + 1;
                ++sp;
                break;
            case GPFO_OP_MES:
                if (a < 0 || a >= GPFO_MAX_SLOTS) return fail(ctx, GPFO_ERR_BAD_PROGRAM, "moment slot out of range");
                if (b < 1 || c < 0 || c + b > nparams) return fail(ctx, GPFO_ERR_BAD_PROGRAM, "MES sample range outside params");
                if (a + 1 > nslots) nslots = a + 1;
                ++sp;
                break;
            case GPFO_OP_HVPOI:
                if (hv) return fail(ctx, GPFO_ERR_UNSUPPORTED, "at most one HVPoI leaf per program");
                if (a < 0 || b < 2 || b > 6 || a + b > GPFO_MAX_SLOTS)
                    return fail(ctx, GPFO_ERR_BAD_PROGRAM, "HVPoI needs 2..6 objectives in valid slots");
                if (a + b > nslots) nslots = a + b;
                hv = 1; ++sp;
                break;
            case GPFO_OP_SUM: case GPFO_OP_PROD:
                if (a < 1 || a > sp) return fail(ctx, GPFO_ERR_BAD_PROGRAM, "aggregation pops more than the stack holds");
                sp -= a - 1;
                break;
            case GPFO_OP_SCALE:
                if (sp < 1 || c < 0 || c >= nparams) return fail(ctx, GPFO_ERR_BAD_PROGRAM, "bad SCALE");
                break;
            default:
                return fail(ctx, GPFO_ERR_BAD_PROGRAM, "unknown opcode");
        }
        if (sp > 8) return fail(ctx, GPFO_ERR_BAD_PROGRAM, "program stack deeper than 8");
    }
    if (sp != 1) return fail(ctx, GPFO_ERR_BAD_PROGRAM, "program must leave exactly one value");
    p.nops = nops; p.nslots = nslots; p.has_hvpoi = hv;
    memcpy(p.ops, ops, sizeof(int32_t) * 4 * nops);
    if (nparams) memcpy(p.params, params, sizeof(double) * nparams);
    ctx->hprog = p;
    CK(cudaMemcpyAsync(ctx->dprog, &ctx->hprog, sizeof(ProgramDev), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->prog_set = true;
    return GPFO_OK;
}

int gpfo_cells_set(gpfo_ctx* ctx, const double* pf_ext, int n_ext_rows, int R, const int32_t* lb, const int32_t* ub,
                   int64_t C) {
    if (!ctx) return GPFO_ERR_STATE;
    cudaSetDevice(ctx->device);
    if (!pf_ext || n_ext_rows < 2 || R < 2 || R > 6 || C < 0 || (C > 0 && (!lb || !ub)))
        return fail(ctx, GPFO_ERR_BAD_SHAPE, "bad cell tables");
    if (gpfo_hvpoi_smem_bytes(n_ext_rows, R) > 227 * 1024)
        return fail(ctx, GPFO_ERR_UNSUPPORTED, "Pareto front too large for the shared-memory Phi table");
    for (int64_t i = 0; i < C * R; ++i)
        if (lb[i] < 0 || lb[i] >= n_ext_rows || ub[i] < 0 || ub[i] >= n_ext_rows)
            return fail(ctx, GPFO_ERR_BAD_SHAPE, "cell index outside pf_ext");
    ++ctx->epoch;
    cudaFree(ctx->dpf); cudaFree(ctx->dlb); cudaFree(ctx->dub); cudaFree(ctx->dcoff); cudaFree(ctx->dcbnd);
    ctx->dpf = nullptr; ctx->dlb = ctx->dub = ctx->dcoff = nullptr; ctx->dcbnd = nullptr; ctx->cells_set = false;
    CK(cudaMalloc((void**)&ctx->dpf, (size_t)n_ext_rows * R * sizeof(double)));
    CK(cudaMalloc((void**)&ctx->dlb, (size_t)(C > 0 ? C : 1) * R * sizeof(int32_t)));
    CK(cudaMalloc((void**)&ctx->dub, (size_t)(C > 0 ? C : 1) * R * sizeof(int32_t)));
    CK(cudaMemcpy(ctx->dpf, pf_ext, (size_t)n_ext_rows * R * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMalloc((void**)&ctx->dcoff, (size_t)(C > 0 ? C : 1) * R * 2 * sizeof(int32_t)));
    CK(cudaMalloc((void**)&ctx->dcbnd, (size_t)(C > 0 ? C : 1) * R * 2 * sizeof(double)));
    if (C > 0) {
        CK(cudaMemcpy(ctx->dlb, lb, (size_t)C * R * sizeof(int32_t), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->dub, ub, (size_t)C * R * sizeof(int32_t), cudaMemcpyHostToDevice));
        // per-cell records of the candidate-per-lane kernel: Phi-table entries and bound values, upper first
        std::vector<int32_t> off((size_t)C * R * 2);
        std::vector<double> bnd((size_t)C * R * 2);
        for (int64_t c = 0; c < C; ++c)
            for (int j = 0; j < R; ++j) {
                const int iu = ub[c * R + j], il = lb[c * R + j];
                off[(c * R + j) * 2] = iu * R + j;
                off[(c * R + j) * 2 + 1] = il * R + j;
                bnd[(c * R + j) * 2] = pf_ext[(size_t)iu * R + j];
                bnd[(c * R + j) * 2 + 1] = pf_ext[(size_t)il * R + j];
            }
        CK(cudaMemcpy(ctx->dcoff, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->dcbnd, bnd.data(), bnd.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    ctx->cells.n_ext = n_ext_rows; ctx->cells.R = R; ctx->cells.C = C;
    ctx->cells.pf_ext = ctx->dpf; ctx->cells.lb = ctx->dlb; ctx->cells.ub = ctx->dub;
    ctx->cells.off = ctx->dcoff; ctx->cells.bnd = ctx->dcbnd;
    ctx->cells_set = true;
    return GPFO_OK;
}

// Stages candidates on the device (if needed) and makes sure scratch buffers are large enough.
static int stage_candidates(gpfo_ctx* ctx, const double* X, int64_t N, int D, const double** dX) {
    if (is_device_ptr(X)) { *dX = X; ctx->tm.h2d_ms = 0.0; return GPFO_OK; }
    NvtxRange nvtx_upload("gpfo:upload");
    CK(ensure(&ctx->dXc, &ctx->cap_xc, (size_t)N * D));
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    const size_t bytes = (size_t)N * D * sizeof(double);
    const void* src = X;
    if (bytes <= GPFO_STAGE_AREA) { memcpy(ctx->hstage, X, bytes); src = ctx->hstage; }
    CK(cudaMemcpyAsync(ctx->dXc, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    *dX = ctx->dXc;
    return GPFO_OK;
}

static int ensure_parts(gpfo_ctx* ctx, int n) {
    if (ctx->cap_parts >= n) return GPFO_OK;
    ++g_alloc_epoch;
    cudaFree(ctx->dpart_val); cudaFree(ctx->dpart_idx);
    ctx->dpart_val = nullptr; ctx->dpart_idx = nullptr; ctx->cap_parts = 0;
    CK(cudaMalloc((void**)&ctx->dpart_val, (size_t)n * sizeof(double)));
    CK(cudaMalloc((void**)&ctx->dpart_idx, (size_t)n * sizeof(long long)));
    ctx->cap_parts = n;
    return GPFO_OK;
}

// What one gpfo_eval call launches, decided once up front.
struct EvalPlan {
    int64_t N;
    int D, variant, grid, hv_grid, pg_grid, slot_first;
    bool use_cluster, hv, want_grad, need_mom, generate;
    bool thin = false;              // latency path (launch_thin)
    double* best_direct = nullptr;  // latency path: where a single-block finish kernel leaves the (value, index) record
};

// Two device buffers of the same size that grow together ([slot][N] moment / coefficient arrays).
static int ensure_pair(gpfo_ctx* ctx, double** a, double** b, size_t* cap, size_t need) {
    if (*cap >= need) return GPFO_OK;
    ++g_alloc_epoch;
    cudaFree(*a); cudaFree(*b);
    *a = *b = nullptr; *cap = 0;
    CK(cudaMalloc((void**)a, need * sizeof(double)));
    CK(cudaMalloc((void**)b, need * sizeof(double)));
    *cap = need;
    return GPFO_OK;
}

// State checks of gpfo_eval: program, models, moment slots, cell tables.
static int check_eval_state(gpfo_ctx* ctx, int D, bool want_grad) {
    if (!ctx->prog_set) return fail(ctx, GPFO_ERR_STATE, "gpfo_eval before gpfo_program_set");
    if (ctx->ngps < 1) return fail(ctx, GPFO_ERR_STATE, "gpfo_eval before gpfo_gp_set");
    bool have[GPFO_MAX_SLOTS] = {false};               // which slots are produced by the registered GPs?
    for (int g = 0; g < ctx->ngps; ++g) {
        const GpHost& gp = ctx->gps[g];
        if (!gp.valid) return fail(ctx, GPFO_ERR_STATE, "a GP below gp count was never set");
        if (gp.D != D) return fail(ctx, GPFO_ERR_BAD_SHAPE, "candidate dimension does not match the models");
        for (int r = 0; r < gp.R; ++r) have[gp.slot0 + r] = true;
    }
    for (int s = 0; s < ctx->hprog.nslots; ++s)
        if (!have[s]) return fail(ctx, GPFO_ERR_STATE, "program references a moment slot no GP produces");
    if (ctx->hprog.has_hvpoi && !ctx->cells_set) return fail(ctx, GPFO_ERR_STATE, "HVPoI program without gpfo_cells_set");
    if (want_grad && ctx->hprog.has_hvpoi &&
        gpfo_hvpoi_smem_bytes_grad(ctx->cells.n_ext, ctx->cells.R) > 227 * 1024)
        return fail(ctx, GPFO_ERR_UNSUPPORTED, "Pareto front too large for the gradient HVPoI kernel");
    return GPFO_OK;
}

// Large gradient batch: forward launches park A = L^-1 Kx, backward launches contract it with the reversed transposed
// stream (2 M^2 flop per candidate and GP in total instead of 3 M^2).  A lives in a bounded scratch, so the batch is
// processed in slabs of whole tiles: forward (all GPs) -> program partials -> backward (all GPs).
static int launch_stored_a(gpfo_ctx* ctx, const EvalPlan& pl, EvalIO io, int* launches_out, int* nparts_out, float* hot_ms_out) {
    cudaStream_t st = ctx->stream;
    const int64_t N = pl.N;
    const int D = pl.D, slot_first = pl.slot_first;
    const bool hv = pl.hv;
    int launches = 0, nparts = 0, rc = GPFO_OK;
    float hot_ms = 0.f;
    const int Tt = gpfo_contract_bwd_tile();
    const int64_t ntiles = (N + Tt - 1) / Tt;
    size_t per_tile = 0;                                   // doubles of A per tile, all GPs
    for (int g = 0; g < ctx->ngps; ++g) per_tile += (size_t)gpfo_contract_astore_stride(ctx->gps[g].Mp / 8);
    // scratch budget: a quarter of the free memory, at most 8 GB.  The driver is asked only when the scratch at hand does not
    // already hold the whole call (cudaMemGetInfo sits between the timing events and takes milliseconds on a busy host).
    size_t budget = ctx->cap_a_scratch * sizeof(double);
    const char* mb = getenv("GPFO_ASTORE_BUDGET_MB");
    if (mb) budget = (size_t)atoll(mb) << 20;                                                    // tests: force several slabs
    else if (budget < (size_t)ntiles * per_tile * sizeof(double)) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        budget = (free_b + ctx->cap_a_scratch * sizeof(double)) / 4;
        if (budget > ((size_t)8 << 30)) budget = (size_t)8 << 30;
    }
    int64_t slab_tiles = (int64_t)(budget / (per_tile * sizeof(double)));
    if (slab_tiles < ctx->num_sms) slab_tiles = ctx->num_sms;
    if (slab_tiles > ntiles) slab_tiles = ntiles;
    CK(ensure(&ctx->da_scratch, &ctx->cap_a_scratch, (size_t)slab_tiles * per_tile));
    const int64_t nslabs = (ntiles + slab_tiles - 1) / slab_tiles;
    const int sgrid = (int)(slab_tiles < ctx->num_sms ? slab_tiles : ctx->num_sms);
    const int64_t slab_n = slab_tiles * Tt;
    const int ppg = hv ? gpfo_hvpoi_grid(slab_n, ctx->num_sms, ctx->cells.n_ext, ctx->cells.R, 1) : gpfo_program_grad_grid(slab_n, ctx->num_sms);
    rc = ensure_parts(ctx, (int)(nslabs * ppg));
    if (rc != GPFO_OK) return rc;
    io.part_val = ctx->dpart_val; io.part_idx = ctx->dpart_idx;
    // forward kernel of the pass: the resident-panel kernel where it beats the 64-candidate tiles (cost model of the
    // large-batch forward), unless a kernel was pinned or the input is too wide for its shared memory
    int mp_max = 0;
    for (int g = 0; g < ctx->ngps; ++g) mp_max = std::max(mp_max, ctx->gps[g].Mp);
    bool v5_forward = ctx->variant == 0 && D <= gpfo_contract5_max_dim(20) && !getenv("GPFO_GRAD_FWD_V12") &&
                      large_batch_cost(20, mp_max) < large_batch_cost(12, mp_max);
    // small training sets: the 32-candidate x 512-row tiling as the forward kernel (A in the 64-candidate layout of round 1's
    // backward kernel), where it beats both
    const bool v21_forward = ctx->variant == 0 && D <= gpfo_contract5_max_dim(21) && !getenv("GPFO_GRAD_FWD_V12") &&
                             large_batch_cost(21, mp_max) < large_batch_cost(12, mp_max) &&
                             large_batch_cost(21, mp_max) < large_batch_cost(20, mp_max);
    if (v21_forward) v5_forward = true;
    // ... and then the backward launch in the same tiling (A in 16-candidate tiles, bulk copies), GPFO_GRAD_BWD_V2=1 keeps round
    // 1's 64-candidate backward kernel (comparisons)
    // (measured, profiles/r02_grad_bench_tall_backward.log: stored-A total 63.1 / 75.2 / 79.9 / 85.1 % of peak at M = 1024 / 2048 /
    // 4096 / 8192 against 62.0 / 70.9 / 74.9 / 75.1 % with the 64-candidate backward kernel)
    static const int bwd5_min_m = getenv("GPFO_BWD5_MIN_M") ? atoi(getenv("GPFO_BWD5_MIN_M")) : 1024;
    const bool v5_backward = v5_forward && !v21_forward && mp_max >= bwd5_min_m && gpfo_contract5_bwd_ok(D) && !getenv("GPFO_GRAD_BWD_V2");
    size_t kx_need = 0;
    for (int g = 0; g < ctx->ngps && !v5_forward; ++g) {
        const size_t cols = ((size_t)ctx->gps[g].Mp + 127) / 128 * 128;
        if (cols * Tt * (size_t)sgrid > kx_need) kx_need = cols * Tt * (size_t)sgrid;
    }
    if (kx_need) CK(ensure(&ctx->dkx_scratch, &ctx->cap_kx_scratch, kx_need));
    nparts = 0;
    for (int64_t k = 0; k < nslabs; ++k) {
        const int64_t n0 = k * slab_n;
        const int64_t ns = (N - n0 < slab_n) ? N - n0 : slab_n;
        const int64_t st_tiles = (ns + Tt - 1) / Tt;
        const int g_s = (int)(st_tiles < sgrid ? st_tiles : sgrid);
        EvalIO sio = io;
        sio.X = io.X + (size_t)n0 * D; sio.N = ns; sio.idx0 = n0;
        sio.mom_mean = io.mom_mean + n0; sio.mom_var = io.mom_var + n0;
        sio.cmu = io.cmu + n0; sio.cva = io.cva + n0;
        if (io.values) sio.values = io.values + n0;
        sio.grads = io.grads + (size_t)n0 * D;
        sio.part_val = io.part_val + nparts; sio.part_idx = io.part_idx + nparts;
        size_t a_off = 0;
        for (int g = 0; g < ctx->ngps; ++g) {
            const GpHost& gh = ctx->gps[g];
            EvalIO gio = sio;
            gio.write_moments = 1; gio.run_program = 0;
            gio.kx_scratch = ctx->dkx_scratch;
            gio.kx_scratch_stride = (long long)(((size_t)gh.Mp + 127) / 128 * 128 * Tt);
            gio.a_scratch = ctx->da_scratch + a_off;
            gio.a_stride = gpfo_contract_astore_stride(gh.Mp / 8);
            a_off += (size_t)slab_tiles * (size_t)gio.a_stride;
            if (k == 0) CK(cudaEventRecord(ctx->gp_ev[g][0], st));
            if (v5_forward) {              // resident-panel forward (no K(X*,X) scratch), same A layout
                rc = ensure_pack(ctx, ctx->gps[g], gpfo_contract_rbc(v21_forward ? 21 : 20), &gio.packed_o, &gio.geom_o);
                if (rc != GPFO_OK) return rc;
                if (v5_backward) gio.a_stride = gpfo_contract5_astore_stride(gh.Mp / 8);
                CK(gpfo_contract5_astore_launch(st, gh.kind, gh.ddev, ctx->dprog, gio, ctx->num_sms, D, v5_backward ? 1 : 0,
                                                v21_forward ? 21 : 20));
            } else {
                CK(gpfo_contract_astore_launch(st, gh.kind, gh.ddev, ctx->dprog, gio, g_s, D));
            }
            if (k == 0) CK(cudaEventRecord(ctx->gp_ev[g][1], st));
            ++launches;
        }
        int parts_k;
        if (hv) {
            EvalIO hio = sio;
            hio.run_program = 1;
            parts_k = gpfo_hvpoi_grid(ns, ctx->num_sms, ctx->cells.n_ext, ctx->cells.R, 1);
            CK(gpfo_hvpoi_launch(st, ctx->dprog, ctx->cells, hio, parts_k, slot_first, 1));
        } else {
            parts_k = gpfo_program_grad_grid(ns, ctx->num_sms);
            CK(gpfo_program_grad_launch(st, ctx->dprog, sio, parts_k));
        }
        ++launches;
        nparts += parts_k;
        a_off = 0;
        for (int g = 0; g < ctx->ngps; ++g) {
            const GpHost& gh = ctx->gps[g];
            EvalIO gio = sio;
            gio.grad_accumulate = g > 0 ? 1 : 0;
            gio.packed_o = gh.dpacked_rt; gio.geom_o = gh.geom_rt;
            gio.a_scratch = ctx->da_scratch + a_off;
            gio.a_stride = gpfo_contract_astore_stride(gh.Mp / 8);
            a_off += (size_t)slab_tiles * (size_t)gio.a_stride;
            if (v5_backward) {
                gio.packed_o = gh.dpacked_rt5; gio.geom_o = gh.geom_rt5;
                gio.a_stride = gpfo_contract5_astore_stride(gh.Mp / 8);
                CK(gpfo_contract5_bwd_launch(st, gh.kind, gh.ddev, gio, ctx->num_sms, D));
            } else
            CK(gpfo_contract_bwd_launch(st, gh.kind, gh.ddev, gio, g_s, D));
            ++launches;
        }
    }
    *launches_out = launches; *nparts_out = nparts; *hot_ms_out = hot_ms;
    return GPFO_OK;
}

// Latency path, a handful of candidates (every GP's forward -- and gradient -- launch would be a row-split one): all contraction
// launches are independent of each other (the gradient launches leave raw sums, the program's partials are applied by the
// finish kernel), so the gradient ones go to a second stream, and ONE finish kernel replaces the per-GP finish kernels, the
// program-gradient kernel and, for up to 32 candidates, the argmax kernel.  Reference call site: the N = 1 L-BFGS-B polish
// evaluations of optim.py:252-276 through acquisition.py:256-257.
// One tile (<= 8 candidates) whose K(X*,X) fits one shared-memory panel: thin.cu's 16-row kernels (any M); beyond, while a row
// split applies, the row-split DMMA kernels (contract2_aux.cu), which amortise their per-chunk work over 128 rows.
static bool thin_rows16(const gpfo_ctx* ctx, int64_t N) {
    if (N > gpfo_thin_max_n()) return false;
    for (int g = 0; g < ctx->ngps; ++g)
        if (ctx->gps[g].Mp > gpfo_thin_panel_cols(N)) return false;     // K(X*,X) of the call must fit one shared-memory panel
    return true;
}
static int thin_ns(const gpfo_ctx* ctx, int64_t N, const PackGeom& geom) {
    return thin_rows16(ctx, N) ? gpfo_thin_nsplit(N, ctx->num_sms, geom.nrb) : split_count(ctx, N, gpfo_split_tile(), geom.nchunks);
}
static bool thin_applies(const gpfo_ctx* ctx, int64_t N, bool hv, bool generate, bool want_grad) {
    static const bool off = getenv("GPFO_NO_THIN") && atoi(getenv("GPFO_NO_THIN")) != 0;
    if (off || ctx->variant != 0 || hv || generate) return false;
    if (thin_rows16(ctx, N)) return true;
    for (int g = 0; g < ctx->ngps; ++g) {
        if (split_count(ctx, N, gpfo_split_tile(), ctx->gps[g].geom_s.nchunks) <= 1) return false;
        if (want_grad && split_count(ctx, N, gpfo_split_tile(), ctx->gps[g].geom_kinv_s.nchunks) <= 1) return false;
    }
    return true;
}
static size_t thin_partials(const gpfo_ctx* ctx, int64_t N, int D, bool want_grad) {
    const size_t ntiles = (size_t)((N + 7) / 8);
    size_t need = 0;
    for (int g = 0; g < ctx->ngps; ++g) {
        need += ntiles * thin_ns(ctx, N, ctx->gps[g].geom_s) * 8 * (1 + RMAX_KERNEL);
        if (want_grad) need += ntiles * thin_ns(ctx, N, ctx->gps[g].geom_kinv_s) * 8 * (size_t)(1 + ctx->gps[g].R) * (1 + D);
    }
    return need;
}
static int launch_thin(gpfo_ctx* ctx, const EvalPlan& pl, EvalIO io, int* launches_out, int* nparts_out) {
    cudaStream_t st = ctx->stream, st2 = ctx->stream2;
    const int64_t N = pl.N;
    const int D = pl.D;
    const size_t ntiles = (size_t)((N + 7) / 8);
    const bool rows16 = thin_rows16(ctx, N);
    ThinFinalArgs a;
    memset(&a, 0, sizeof(a));
    a.ngps = ctx->ngps; a.D = D; a.want_grad = pl.want_grad ? 1 : 0; a.best_out = pl.best_direct;
    int launches = 0;
    size_t off = 0;
    if (pl.want_grad) {
        CK(cudaEventRecord(ctx->ev_fork, st));
        CK(cudaStreamWaitEvent(st2, ctx->ev_fork, 0));
    }
    for (int g = 0; g < ctx->ngps; ++g) {
        const GpHost& gh = ctx->gps[g];
        EvalIO gio = io;
        gio.write_moments = 0; gio.run_program = 0;
        gio.packed_o = gh.dpacked_s; gio.geom_o = gh.geom_s;
        gio.nsplit = thin_ns(ctx, N, gh.geom_s);
        gio.partials = ctx->dthin + off;
        a.gp[g] = gh.ddev; a.part_f[g] = gio.partials; a.ns_f[g] = gio.nsplit;
        a.R[g] = gh.R; a.slot0[g] = gh.slot0;
        off += ntiles * gio.nsplit * 8 * (1 + RMAX_KERNEL);
        if (!ctx->capturing) CK(cudaEventRecord(ctx->gp_ev[g][0], st));
        if (rows16) CK(gpfo_thin_contract_launch(st, gh.kind, gh.ddev, gio, D, 0));
        else CK(gpfo_thin_forward_launch(st, gh.kind, gh.ddev, gio, D));
        if (!ctx->capturing) CK(cudaEventRecord(ctx->gp_ev[g][1], st));
        ++launches;
    }
    if (pl.want_grad) {
        for (int g = 0; g < ctx->ngps; ++g) {
            const GpHost& gh = ctx->gps[g];
            EvalIO gio = io;
            gio.packed_o = gh.dpacked_kinv_s; gio.geom_o = gh.geom_kinv_s;
            gio.nsplit = thin_ns(ctx, N, gh.geom_kinv_s);
            gio.partials = ctx->dthin + off;
            a.part_g[g] = gio.partials; a.ns_g[g] = gio.nsplit;
            off += ntiles * gio.nsplit * 8 * (size_t)(1 + gh.R) * (1 + D);
            if (rows16) CK(gpfo_thin_contract_launch(st2, gh.kind, gh.ddev, gio, D, 1));
            else CK(gpfo_thin_grad_launch(st2, gh.kind, gh.ddev, gio, D));
            ++launches;
        }
        CK(cudaEventRecord(ctx->ev_join, st2));
        CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    }
    CK(gpfo_thin_final_launch(st, ctx->dprog, io, a));
    ++launches;
    *launches_out = launches; *nparts_out = gpfo_thin_final_grid(N);
    return GPFO_OK;
}

// Everything else: one forward launch per GP (row-split for few candidates), then the HVPoI / program-gradient kernel and,
// for gradients, the dense K^-1 pass per GP.
static int launch_standard(gpfo_ctx* ctx, const EvalPlan& pl, EvalIO io, int* launches_out, int* nparts_out, float* hot_ms_out) {
    cudaStream_t st = ctx->stream;
    const int64_t N = pl.N;
    const int D = pl.D, variant = pl.variant, grid = pl.grid, hv_grid = pl.hv_grid, pg_grid = pl.pg_grid, slot_first = pl.slot_first;
    const bool hv = pl.hv, want_grad = pl.want_grad, need_mom = pl.need_mom, generate = pl.generate, use_cluster = pl.use_cluster;
    int launches = 0, nparts = grid, fwd_parts = grid;
    float hot_ms = 0.f;
    for (int g = 0; g < ctx->ngps; ++g) {
        EvalIO gio = io;
        gio.write_moments = need_mom ? 1 : 0;
        gio.run_program = (!hv && !want_grad && g == ctx->ngps - 1) ? 1 : 0;
        if (variant == 11 || variant == 12 || variant == 13 || variant == 14 || (variant >= 120 && variant <= 129)) {   // 12x: profiling builds of 12
            // per-CTA K(X*,X) scratch: one slot per (training point, candidate of the tile), rounded up to whole groups
            const size_t cols = ((size_t)ctx->gps[g].Mp + 127) / 128 * 128;
            const size_t stride = cols * (size_t)gpfo_contract_tile(variant);
            CK(ensure(&ctx->dkx_scratch, &ctx->cap_kx_scratch, stride * (size_t)grid));
            gio.kx_scratch = ctx->dkx_scratch;
            if (variant == 127) {          // profiling build: phase counters land at the start of the values buffer
                CK(ensure(&ctx->dvalues, &ctx->cap_values, (size_t)(N > 65536 ? N : 65536)));
                gio.debug = reinterpret_cast<long long*>(ctx->dvalues);
                gio.values = nullptr;
            }
            gio.kx_scratch_stride = (long long)stride;
        }
        if (variant >= 19 && variant <= 39) {   // v5 reads its block stream and layout from the launch record
            gio.packed_o = ctx->gps[g].hdev.packed; gio.geom_o = ctx->gps[g].hdev.geom;
            if (variant == 27 || variant == 28) {           // profiling build: phase counters land at the start of the values buffer
                CK(ensure(&ctx->dvalues, &ctx->cap_values, (size_t)(N > 65536 ? N : 65536)));
                gio.debug = reinterpret_cast<long long*>(ctx->dvalues);
                gio.values = nullptr;
            }
        }
        if (!ctx->capturing) CK(cudaEventRecord(ctx->gp_ev[g][0], st));
        const int ns = generate ? 1 : split_count(ctx, N, gpfo_split_tile(), ctx->gps[g].geom_s.nchunks);
        if (ns > 1) {
            gio.packed_o = ctx->gps[g].dpacked_s; gio.geom_o = ctx->gps[g].geom_s;
            gio.nsplit = ns; gio.partials = ctx->dpartials;
            CK(gpfo_contract_split_launch(st, ctx->gps[g].kind, ctx->gps[g].ddev, ctx->dprog, gio, D));
            ++launches;
            fwd_parts = (int)((N + 127) / 128);
        } else {
#ifdef GPFO_BUILD_EXPERIMENTS
            if (use_cluster) CK(gpfo_contract3_launch(st, ctx->gps[g].kind, ctx->gps[g].ddev, ctx->dprog, gio, grid, D));
            else
#endif
            CK(gpfo_contract_launch(st, variant, ctx->gps[g].kind, ctx->gps[g].ddev, ctx->dprog, gio, grid, D));
            fwd_parts = grid;
        }
        if (!ctx->capturing) CK(cudaEventRecord(ctx->gp_ev[g][1], st));   // one event pair per GP: nothing waits on the host between launches
        ++launches;
    }
    nparts = fwd_parts;
    if (hv) {
        EvalIO hio = io;
        hio.run_program = 1;
        CK(gpfo_hvpoi_launch(st, ctx->dprog, ctx->cells, hio, hv_grid, slot_first, want_grad ? 1 : 0));
        ++launches;
        nparts = hv_grid;
    } else if (want_grad) {
        CK(gpfo_program_grad_launch(st, ctx->dprog, io, pg_grid));
        ++launches;
        nparts = pg_grid;
    }
    if (want_grad) {
        // chain rule through each GP: W = K^-1 Kx on DMMA, then the (1+D)-wide reduction over the training points
        const int ggrid = gpfo_contract_grad_grid(N, ctx->num_sms);
        for (int g = 0; g < ctx->ngps; ++g) {
            EvalIO gio = io;
            gio.grad_accumulate = g > 0 ? 1 : 0;
            const GpHost& gh = ctx->gps[g];
            const int ns_small = split_count(ctx, N, gpfo_split_tile(), gh.geom_kinv_s.nchunks);
            const int ns_wide = ns_small > 1 ? 1 : grad_split_wide(ctx, N, gh.hdev.geom_kinv.nchunks);
            if (ns_small > 1 || ns_wide > 1) {
                gio.partials = ctx->dpartials;
                gio.nsplit = ns_small > 1 ? ns_small : ns_wide;
                gio.packed_o = ns_small > 1 ? gh.dpacked_kinv_s : gh.dpacked_kinv;
                gio.geom_o = ns_small > 1 ? gh.geom_kinv_s : gh.hdev.geom_kinv;
                CK(gpfo_contract_grad_split_launch(st, gh.kind, gh.ddev, gio, D, ns_small > 1 ? 0 : 1));
                ++launches;
            } else {
                CK(gpfo_contract_grad_launch(st, gh.kind, gh.ddev, gio, ggrid, D));
            }
            ++launches;
        }
    }

    *launches_out = launches; *nparts_out = nparts; *hot_ms_out = hot_ms;
    return GPFO_OK;
}

int gpfo_eval(gpfo_ctx* ctx, const double* Xcand, int64_t N, int D, double* out_values, double* out_grads,
              double* best_value, int64_t* best_index) {
    if (!ctx) return GPFO_ERR_STATE;
    cudaSetDevice(ctx->device);
    if (N < 0) return fail(ctx, GPFO_ERR_BAD_SHAPE, "N < 0");
    const bool want_grad = out_grads != nullptr;
    int rc = check_eval_state(ctx, D, want_grad);
    if (rc != GPFO_OK) return rc;
    if (N == 0) {
        if (best_index) *best_index = -1;
        if (best_value) *best_value = 0.0;
        memset(&ctx->tm.h2d_ms, 0, sizeof(double) * 4);
        ctx->tm.launches = 0;
        return GPFO_OK;
    }
    const bool generate = ctx->gen_active;
    if (!Xcand && !generate) return fail(ctx, GPFO_ERR_BAD_SHAPE, "null candidates");
    if (generate && want_grad) return fail(ctx, GPFO_ERR_UNSUPPORTED, "gradients of generated candidates are not supported");
    cudaStream_t st = ctx->stream;
    // candidates: device pointers are read in place; host rows are copied to the device right before the launches (below)
    const bool host_in = !generate && !is_device_ptr(Xcand);
    const double* dX = generate ? nullptr : Xcand;
    const size_t xbytes = (size_t)N * D * sizeof(double);
    const bool stage_x = host_in && xbytes <= GPFO_STAGE_AREA;
    if (host_in) { CK(ensure(&ctx->dXc, &ctx->cap_xc, (size_t)N * D)); dX = ctx->dXc; }
    ctx->tm.h2d_ms = 0.0;

    int mp_max = 0;
    for (int g = 0; g < ctx->ngps; ++g) mp_max = std::max(mp_max, ctx->gps[g].Mp);
    int variant = effective_variant(ctx, N, mp_max);
    if (generate && variant == 10) variant = 12;
    if (variant == 19 && D > gpfo_contract5_max_dim(20)) variant = 6;
    if ((variant == 20 || variant == 21) && D > gpfo_contract5_max_dim(variant)) variant = 12;   // very wide inputs exceed v5's shared memory
    // gradients with a resident-panel variant: a large batch (or a pinned variant) means the stored-A pass, which is entered as
    // variant 12 (it picks its own forward kernel); a mid-size batch below the stored-A threshold keeps the variant as the
    // forward launch of the dense K^-1 pass
    const bool v5_choice = variant == 20 || variant == 21;
    if (v5_choice && want_grad && (ctx->variant != 0 || N >= (int64_t)64 * 4 * ctx->num_sms)) variant = 12;
    // Gradient batches from about 40 candidates per SM up take the stored-A pass (2 M^2 flop) as well: its forward launch walks
    // 16-candidate tiles (all SMs busy from N = 16 SMs), and from there on even a partly filled wave of 64-candidate backward
    // tiles costs less than the dense K^-1 pass (3 M^2 flop); measured in profiles/r02_grad_mid_stored_a.log.
    // GPFO_ASTORE_MIN_N moves the threshold.
    if (want_grad && ctx->variant == 0 && variant != 12) {      // (variant 12 from the wave model already means stored-A)
        const char* e = getenv("GPFO_ASTORE_MIN_N");
        // both launches of the pass on 16-candidate tiles (tall-chunk kernels): worth it from one wave of such tiles
        // (profiles/r02_grad_mid_stored_a.log: x1.0-1.6 over the dense pass for N = 2.4e3 ... 4.7e3, x1.3-1.8 from 1.2e4)
        const bool tall = mp_max >= 1024 && D <= gpfo_contract5_max_dim(20) && gpfo_contract5_bwd_ok(D) &&
                          large_batch_cost(20, mp_max) < large_batch_cost(12, mp_max);
        const int64_t min_n = e ? atoll(e) : (int64_t)(tall ? 16 : 40) * ctx->num_sms;
        if (N >= min_n && D <= gpfo_contract5_max_dim(20)) variant = 12;
    }
    const bool use_cluster = variant == 10;
    const int rbc = gpfo_contract_rbc(variant);
    for (int g = 0; g < ctx->ngps; ++g) {
#ifdef GPFO_BUILD_EXPERIMENTS
        if (use_cluster) rc = ensure_cluster_pack(ctx, ctx->gps[g]);
        else
#endif
        if (ctx->gps[g].packed_rbc != rbc) rc = repack(ctx, ctx->gps[g], rbc);
        if (rc != GPFO_OK) return rc;
    }

    const bool hv = ctx->hprog.has_hvpoi != 0;
    const bool need_mom = hv || ctx->ngps > 1 || want_grad;
    if (want_grad)
        for (int g = 0; g < ctx->ngps; ++g) { rc = ensure_kinv(ctx, ctx->gps[g]); if (rc != GPFO_OK) return rc; }
    // large gradient batches take the stored-A path (GPFO_GRAD_DENSE=1 keeps the dense K^-1 pass, for comparisons)
    const char* env_dense = getenv("GPFO_GRAD_DENSE");
    const bool astore = want_grad && variant == 12 && gpfo_contract_bwd_smem(D) <= (size_t)227 * 1024 &&
                        !(env_dense && atoi(env_dense) != 0);
    if (astore)
        for (int g = 0; g < ctx->ngps; ++g) { rc = ensure_rt(ctx, ctx->gps[g]); if (rc != GPFO_OK) return rc; }
    EvalIO io;
    memset(&io, 0, sizeof(io));
    io.X = dX; io.N = N; io.mstride = N;
    if (generate) {
        io.gen_A = ctx->dgen; io.gen_b = ctx->dgen + GPFO_MAX_DIM;
        io.gen_seed = ctx->gen_seed; io.gen_first = ctx->gen_first;
    }
    if (need_mom) {
        rc = ensure_pair(ctx, &ctx->dmom_mean, &ctx->dmom_var, &ctx->cap_mom, (size_t)GPFO_MAX_SLOTS * N);
        if (rc != GPFO_OK) return rc;
        io.mom_mean = ctx->dmom_mean; io.mom_var = ctx->dmom_var;
    }
    const bool dev_out = out_values && is_device_ptr(out_values);
    if (out_values) {
        if (dev_out) io.values = out_values;
        else { CK(ensure(&ctx->dvalues, &ctx->cap_values, (size_t)N)); io.values = ctx->dvalues; }
    }
    const bool dev_grads = want_grad && is_device_ptr(out_grads);
    if (want_grad) {
        rc = ensure_pair(ctx, &ctx->dcmu, &ctx->dcva, &ctx->cap_coef, (size_t)GPFO_MAX_SLOTS * N);
        if (rc != GPFO_OK) return rc;
        io.cmu = ctx->dcmu; io.cva = ctx->dcva;
        if (dev_grads) io.grads = out_grads;
        else { CK(ensure(&ctx->dgrads, &ctx->cap_grads, (size_t)N * D)); io.grads = ctx->dgrads; }
    }
#ifdef GPFO_BUILD_EXPERIMENTS
    const int grid = use_cluster ? gpfo_contract3_clusters(N, ctx->num_sms) : gpfo_contract_grid(variant, N, ctx->num_sms);
#else
    const int grid = gpfo_contract_grid(variant, N, ctx->num_sms);
#endif
    const int hv_grid = hv ? gpfo_hvpoi_grid(N, ctx->num_sms, ctx->cells.n_ext, ctx->cells.R, want_grad ? 1 : 0) : 1;
    const int pg_grid = gpfo_program_grad_grid(N, ctx->num_sms);
    const bool thin = !astore && thin_applies(ctx, N, hv, generate, want_grad);
    int max_parts = grid > hv_grid ? grid : hv_grid;
    if (pg_grid > max_parts) max_parts = pg_grid;
    if (thin && gpfo_thin_final_grid(N) > max_parts) max_parts = gpfo_thin_final_grid(N);
    if (max_parts < ctx->num_sms) max_parts = ctx->num_sms;      // row-split finish kernel: <= num_sms / 2 * 8 / 128 blocks
    rc = ensure_parts(ctx, max_parts);
    if (rc != GPFO_OK) return rc;
    io.part_val = ctx->dpart_val; io.part_idx = ctx->dpart_idx;

    int launches = 0;
    float hot_ms = 0.f;
    rc = ensure_partials(ctx);
    if (rc != GPFO_OK) return rc;
    int nparts = grid;
    int slot_first = 0;
    if (hv)
        for (int i = 0; i < ctx->hprog.nops; ++i)
            if (ctx->hprog.ops[4 * i] == GPFO_OP_HVPOI) {
                slot_first = ctx->hprog.ops[4 * i + 1];
                if (ctx->hprog.ops[4 * i + 2] != ctx->cells.R)
                    return fail(ctx, GPFO_ERR_BAD_SHAPE, "HVPoI objective count does not match the cell tables");
            }
    EvalPlan pl;
    pl.N = N; pl.D = D; pl.variant = variant; pl.grid = grid; pl.hv_grid = hv_grid; pl.pg_grid = pg_grid;
    pl.slot_first = slot_first; pl.use_cluster = use_cluster; pl.hv = hv; pl.want_grad = want_grad;
    pl.need_mom = need_mom; pl.generate = generate;
    const size_t vbytes = (size_t)N * sizeof(double), gbytes = (size_t)N * D * sizeof(double);
    char* const hv_stage = ctx->hstage + GPFO_STAGE_AREA, *const hg_stage = ctx->hstage + 2 * GPFO_STAGE_AREA;
    char* const hb_stage = ctx->hstage + 3 * GPFO_STAGE_AREA;
    const bool stage_v = out_values && !dev_out && vbytes <= GPFO_STAGE_AREA;
    const bool stage_g = want_grad && !dev_grads && gbytes <= GPFO_STAGE_AREA;
    // latency path: the finish kernel writes the results to the pinned (device-mapped) staging areas itself -- no copy nodes
    // behind a call of a few microseconds' worth of kernels (GPFO_THIN_COPIES=1 keeps the copies).  The candidates still
    // arrive by a copy: read in place by every CTA of the contraction launches they cost 9-18 us of serialised PCIe reads
    // (profiles/r02_thin_phase_stamps.log).
    static const bool thin_copies = getenv("GPFO_THIN_COPIES") && atoi(getenv("GPFO_THIN_COPIES")) != 0;
    pl.thin = thin;
    const bool direct = pl.thin && !thin_copies;
    const bool direct_x = false, direct_v = direct && stage_v, direct_g = direct && stage_g;
    const bool direct_b = direct && gpfo_thin_final_grid(N) == 1;
    if (pl.thin) {
        CK(ensure(&ctx->dthin, &ctx->cap_thin, thin_partials(ctx, N, D, want_grad)));
        if (direct_x) io.X = reinterpret_cast<const double*>(ctx->hstage);
        if (direct_v) io.values = reinterpret_cast<double*>(hv_stage);
        if (direct_g) io.grads = reinterpret_cast<double*>(hg_stage);
        if (direct_b) pl.best_direct = reinterpret_cast<double*>(hb_stage);
    }

    // Everything the call puts on the stream, in order: candidate upload, kernels, final argmax, result download.
    // `timed` = with the event records of the timing breakdown (not capturable into a graph).
    auto enqueue = [&](bool timed) -> int {
        if (host_in && !direct_x) {
            NvtxRange nvtx_upload("gpfo:upload");
            if (timed) CK(cudaEventRecord(ctx->ev[2], st));
            CK(cudaMemcpyAsync(ctx->dXc, stage_x ? (const void*)ctx->hstage : (const void*)Xcand, xbytes, cudaMemcpyHostToDevice, st));
            if (timed) CK(cudaEventRecord(ctx->ev[3], st));
        }
        if (timed) CK(cudaEventRecord(ctx->ev[4], st));
        {
            NvtxRange nvtx_eval("gpfo:eval");
            const int rc2 = pl.thin ? launch_thin(ctx, pl, io, &launches, &nparts)
                            : astore ? launch_stored_a(ctx, pl, io, &launches, &nparts, &hot_ms)
                                     : launch_standard(ctx, pl, io, &launches, &nparts, &hot_ms);
            if (rc2 != GPFO_OK) return rc2;
        }
        if (!direct_b) {
            NvtxRange nvtx_argmax("gpfo:argmax");
            gpfo_argmax_final(st, ctx->dpart_val, ctx->dpart_idx, nparts, ctx->dbest_val, ctx->dbest_idx);
            CK(cudaGetLastError());
            ++launches;
        }
        if (timed) CK(cudaEventRecord(ctx->ev[5], st));
        NvtxRange nvtx_download("gpfo:download");
        if (out_values && !dev_out && !direct_v)
            CK(cudaMemcpyAsync(stage_v ? (void*)hv_stage : (void*)out_values, ctx->dvalues, vbytes, cudaMemcpyDeviceToHost, st));
        if (want_grad && !dev_grads && !direct_g)
            CK(cudaMemcpyAsync(stage_g ? (void*)hg_stage : (void*)out_grads, ctx->dgrads, gbytes, cudaMemcpyDeviceToHost, st));
        if (!direct_b) CK(cudaMemcpyAsync(hb_stage, ctx->dbest_val, 16, cudaMemcpyDeviceToHost, st));
        if (timed) CK(cudaEventRecord(ctx->ev[1], st));
        return GPFO_OK;
    };

    // ---- small calls: replay a captured graph of the sequence above (shape seen before, nothing it depends on changed)
    static const bool graphs_off = getenv("GPFO_NO_GRAPHS") && atoi(getenv("GPFO_NO_GRAPHS")) != 0;
    const bool graphable = !graphs_off && !generate && host_in && stage_x && !astore && N <= 1024 &&
                           (!out_values || stage_v) && (!want_grad || stage_g);
    bool replayed = false;
    if (stage_x) memcpy(ctx->hstage, Xcand, xbytes);
    if (graphable) {
        const unsigned long long key[4] = {(unsigned long long)N, (unsigned long long)D | (want_grad ? 256ull : 0ull) | (out_values ? 512ull : 0ull),
                                           ctx->epoch, g_alloc_epoch};
        gpfo_ctx::GraphEntry* hit = nullptr;
        for (auto& g : ctx->graphs)
            if (g.key[0] == key[0] && g.key[1] == key[1] && g.key[2] == key[2] && g.key[3] == key[3]) { hit = &g; break; }
        if (hit && !hit->exec) {                           // second sighting: capture and instantiate
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
                ctx->capturing = true;
                const int rc2 = enqueue(false);
                ctx->capturing = false;
                const cudaError_t e = cudaStreamEndCapture(st, &graph);
                if (rc2 == GPFO_OK && e == cudaSuccess && graph &&
                    cudaGraphInstantiate(&hit->exec, graph, 0) == cudaSuccess) {
                    hit->launches = launches; hit->variant = variant;
                } else {
                    hit->exec = nullptr;
                    cudaGetLastError();
                }
                if (graph) cudaGraphDestroy(graph);
                launches = 0;
            }
        }
        if (hit && hit->exec) {
            CK(cudaEventRecord(ctx->ev[4], st));
            CK(cudaGraphLaunch(hit->exec, st));
            CK(cudaEventRecord(ctx->ev[1], st));
            launches = hit->launches;
            replayed = true;
            ++ctx->graph_replays;
        } else if (!hit) {
            if (ctx->graphs.size() >= 32) drop_graphs(ctx);
            gpfo_ctx::GraphEntry ge;
            memcpy(ge.key, key, sizeof(key)); ge.exec = nullptr; ge.launches = 0; ge.variant = variant;
            ctx->graphs.push_back(ge);
        }
    }
    if (!replayed) {
        rc = enqueue(true);
        if (rc != GPFO_OK) return rc;
    }
    CK(cudaStreamSynchronize(st));
    if (stage_v) memcpy(out_values, hv_stage, vbytes);
    if (stage_g) memcpy(out_grads, hg_stage, gbytes);
    double hbest;
    long long hidx;
    memcpy(&hbest, hb_stage, sizeof(double));
    memcpy(&hidx, hb_stage + 8, sizeof(long long));
    if (best_value) *best_value = hbest;
    if (best_index) *best_index = (int64_t)hidx;
    if (replayed) {                                        // one event pair around the whole graph: copies and kernels together
        float gms = 0.f;
        cudaEventElapsedTime(&gms, ctx->ev[4], ctx->ev[1]);
        ctx->tm.h2d_ms = 0.0; ctx->tm.eval_ms = gms; ctx->tm.d2h_ms = 0.0; ctx->tm.hot_kernel_ms = 0.0;
        ctx->tm.launches = launches;
        ctx->tm.variant = variant;
        return GPFO_OK;
    }
    float ms = 0.f;
    if (host_in && !direct_x) { cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); ctx->tm.h2d_ms = ms; }
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); ctx->tm.eval_ms = ms;
    cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[1]); ctx->tm.d2h_ms = ms;
    hot_ms = 0.f;
    for (int g = 0; g < ctx->ngps; ++g) { cudaEventElapsedTime(&ms, ctx->gp_ev[g][0], ctx->gp_ev[g][1]); hot_ms += ms; }
    ctx->tm.hot_kernel_ms = hot_ms;
    ctx->tm.launches = launches;
    ctx->tm.variant = variant;
    return GPFO_OK;
}

// affine map of design.py:55-70 for generative_domain = unit cube: A = upper - lower, b = -1 * A + upper
static void gen_map(int D, const double* lower, const double* upper, double* A, double* b) {
    for (int d = 0; d < D; ++d) {
        A[d] = (upper[d] - lower[d]) / (1.0 - 0.0);
        b[d] = -1.0 * A[d] + upper[d];
    }
}

int gpfo_eval_random(gpfo_ctx* ctx, uint64_t seed, int64_t first_row, int64_t N, int D, const double* lower,
                     const double* upper, double* out_values, double* best_value, int64_t* best_index) {
    if (!ctx) return GPFO_ERR_STATE;
    if (D < 1 || D > GPFO_MAX_DIM || !lower || !upper || first_row < 0) return fail(ctx, GPFO_ERR_BAD_SHAPE, "bad generator arguments");
    cudaSetDevice(ctx->device);
    double host[2 * GPFO_MAX_DIM] = {0.0};
    gen_map(D, lower, upper, host, host + GPFO_MAX_DIM);
    if (!ctx->dgen) CK(cudaMalloc((void**)&ctx->dgen, sizeof(host)));
    CK(cudaMemcpy(ctx->dgen, host, sizeof(host), cudaMemcpyHostToDevice));
    ctx->gen_active = true; ctx->gen_seed = seed; ctx->gen_first = first_row;
    const int rc = gpfo_eval(ctx, nullptr, N, D, out_values, nullptr, best_value, best_index);
    ctx->gen_active = false;
    return rc;
}

int gpfo_random_rows(uint64_t seed, int64_t first_row, int64_t n, int D, const double* lower, const double* upper,
                     double* out) {
    if (D < 1 || D > GPFO_MAX_DIM || !lower || !upper || !out || n < 0 || first_row < 0) return GPFO_ERR_BAD_SHAPE;
    double A[GPFO_MAX_DIM], b[GPFO_MAX_DIM];
    gen_map(D, lower, upper, A, b);
    for (int64_t r = 0; r < n; ++r)
        for (int d = 0; d < D; ++d) {
            const double u = gpfo_uniform(seed, (unsigned long long)(first_row + r) * D + d);
            volatile double prod = u * A[d];               // separate roundings, as the device's __dmul_rn / __dadd_rn
            out[r * D + d] = prod + b[d];
        }
    return GPFO_OK;
}

int gpfo_predict(gpfo_ctx* ctx, int gp_id, const double* X, int64_t N, int D, double* mean, double* var) {
    if (!ctx) return GPFO_ERR_STATE;
    cudaSetDevice(ctx->device);
    if (gp_id < 0 || gp_id >= GPFO_MAX_GPS || !ctx->gps[gp_id].valid) return fail(ctx, GPFO_ERR_STATE, "gp not set");
    GpHost& g = ctx->gps[gp_id];
    if (g.D != D) return fail(ctx, GPFO_ERR_BAD_SHAPE, "point dimension does not match the model");
    if (N < 0) return fail(ctx, GPFO_ERR_BAD_SHAPE, "N < 0");
    if (N == 0) return GPFO_OK;
    if (!X) return fail(ctx, GPFO_ERR_BAD_SHAPE, "null points");
    cudaStream_t st = ctx->stream;
    const double* dX = nullptr;
    int rc = stage_candidates(ctx, X, N, D, &dX);
    if (rc != GPFO_OK) return rc;
    int variant = effective_variant(ctx, N, g.Mp);
    if (variant == 19 && D > gpfo_contract5_max_dim(20)) variant = 6;
    if ((variant == 20 || variant == 21) && D > gpfo_contract5_max_dim(variant)) variant = 12;
    if (variant == 10 || variant == 11 || variant == 12 || variant == 13 || variant == 14) variant = (N >= (int64_t)32 * ctx->num_sms) ? 4 : ((N >= (int64_t)16 * ctx->num_sms) ? 5 : 6);
    const int rbc = gpfo_contract_rbc(variant);
    if (g.packed_rbc != rbc) { rc = repack(ctx, g, rbc); if (rc != GPFO_OK) return rc; }
    rc = ensure_pair(ctx, &ctx->dmom_mean, &ctx->dmom_var, &ctx->cap_mom, (size_t)GPFO_MAX_SLOTS * N);
    if (rc != GPFO_OK) return rc;
    const int grid = gpfo_contract_grid(variant, N, ctx->num_sms);
    rc = ensure_parts(ctx, grid);
    if (rc != GPFO_OK) return rc;
    EvalIO io;
    memset(&io, 0, sizeof(io));
    io.X = dX; io.N = N; io.mstride = N; io.mom_mean = ctx->dmom_mean; io.mom_var = ctx->dmom_var;
    io.part_val = ctx->dpart_val; io.part_idx = ctx->dpart_idx;
    io.write_moments = 1; io.run_program = 0;
    if (variant >= 19 && variant <= 39) { io.packed_o = g.hdev.packed; io.geom_o = g.hdev.geom; }
    const int ns = split_count(ctx, N, gpfo_split_tile(), g.geom_s.nchunks);
    if (ns > 1) {
        rc = ensure_partials(ctx);
        if (rc != GPFO_OK) return rc;
        io.packed_o = g.dpacked_s; io.geom_o = g.geom_s; io.nsplit = ns; io.partials = ctx->dpartials;
        CK(gpfo_contract_split_launch(st, g.kind, g.ddev, ctx->dprog, io, D));
    } else {
        CK(gpfo_contract_launch(st, variant, g.kind, g.ddev, ctx->dprog, io, grid, D));
    }
    // [slot][N] -> N x R
    for (int r = 0; r < g.R; ++r) {
        const size_t off = (size_t)(g.slot0 + r) * N;
        if (mean) CK(cudaMemcpy2DAsync(mean + r, (size_t)g.R * sizeof(double), ctx->dmom_mean + off, sizeof(double),
                                       sizeof(double), (size_t)N, cudaMemcpyDefault, st));
        if (var) CK(cudaMemcpy2DAsync(var + r, (size_t)g.R * sizeof(double), ctx->dmom_var + off, sizeof(double),
                                      sizeof(double), (size_t)N, cudaMemcpyDefault, st));
    }
    CK(cudaStreamSynchronize(st));
    return GPFO_OK;
}

int gpfo_timings_get(const gpfo_ctx* ctx, gpfo_timings* out) {
    if (!ctx || !out) return GPFO_ERR_STATE;
    *out = ctx->tm;
    return GPFO_OK;
}

int gpfo_set_variant(gpfo_ctx* ctx, int variant) {
    if (!ctx) return GPFO_ERR_STATE;
#ifdef GPFO_BUILD_EXPERIMENTS
    static const int known[] = {0, 4, 5, 6, 10, 11, 12, 13, 14, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 41, 42, 43, 121, 122, 123, 125, 127};
#else
    static const int known[] = {0, 4, 5, 6, 11, 12, 19, 20, 21};
#endif
    bool ok = false;
    for (int k : known) ok = ok || k == variant;
    if (!ok) return fail(ctx, GPFO_ERR_BAD_SHAPE, "unknown kernel variant");
    ++ctx->epoch;
    ctx->variant = variant;
    return GPFO_OK;
}

int gpfo_select_variant(gpfo_ctx* ctx, int64_t N, int* variant) {
    if (!ctx || !variant || N < 0) return GPFO_ERR_STATE;
    int mp_max = 0;
    for (int g = 0; g < ctx->ngps; ++g) mp_max = std::max(mp_max, ctx->gps[g].Mp);
    int v = effective_variant(ctx, N, mp_max);
    if (v == 20 || v == 21) {
        int D = 0;
        for (int g = 0; g < ctx->ngps; ++g) D = std::max(D, ctx->gps[g].D);
        if (D > gpfo_contract5_max_dim(v)) v = 12;
    }
    *variant = v;
    return GPFO_OK;
}

int gpfo_measure_dmma_peak(gpfo_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return GPFO_ERR_STATE;
    cudaSetDevice(ctx->device);
    const double v = gpfo_dmma_peak(ctx->stream, ctx->num_sms);
    if (v < 0.0) return fail(ctx, GPFO_ERR_CUDA, "DMMA peak kernel failed");
    *tflops = v;
    return GPFO_OK;
}

int gpfo_ndtr_fast_host(const double* z, int64_t n, double* out) {
    if (!z || !out || n < 0) return GPFO_ERR_BAD_SHAPE;
    double tab32[32];
    for (int i = 0; i < 32; ++i) tab32[i] = exp2(i * (1.0 / 32.0));
    for (int64_t i = 0; i < n; ++i) out[i] = gpfo_ndtr_fast(z[i], tab32);
    return GPFO_OK;
}

int gpfo_measure_fp64_pipes(gpfo_ctx* ctx, int mode, double* dmma_tflops, double* dfma_tflops) {
    if (!ctx || !dmma_tflops || !dfma_tflops) return GPFO_ERR_STATE;
    if (mode < 0 || mode > 2) return fail(ctx, GPFO_ERR_BAD_SHAPE, "mode must be 0, 1 or 2");
    cudaSetDevice(ctx->device);
    double out[2] = {0.0, 0.0};
    if (gpfo_fp64_pipes(ctx->stream, ctx->num_sms, mode, out) != 0) return fail(ctx, GPFO_ERR_CUDA, "pipe benchmark failed");
    *dmma_tflops = out[0]; *dfma_tflops = out[1];
    return GPFO_OK;
}

}  // extern "C"
