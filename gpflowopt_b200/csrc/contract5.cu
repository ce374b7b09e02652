// Contraction kernel v5 (resident K(X*,X) panel, no global scratch): launches.  Kernel: contract5_kernel.cuh.
#include "contract5_kernel.cuh"

// variants: 20 = 16 candidates x 1024-row chunks, 21 = 32 candidates x 512-row chunks, 19 = 8 candidates x 1024-row chunks
// (ring: 2 quads of row blocks per warp)
cudaError_t gpfo_contract5_launch(cudaStream_t st, int variant, int kind, const GpDev* d_gp, const ProgramDev* d_prog,
                                  EvalIO io, int grid, int D) {
    // a last chunk that the row blocks do not fill: the instantiation that runs it through the specialised paths (bit 16384)
    const bool short_last = io.geom_o.nrb % io.geom_o.rbc != 0;
    if (variant == 21) {
        if (short_last) return launch5<4, 4, 2, 16384>(st, kind, d_gp, d_prog, io, grid, D);
        return launch5<4, 4, 2>(st, kind, d_gp, d_prog, io, grid, D);
    }
    if (variant == 19) {                          // 8 candidates x 1024-row chunks: small batches (one wave of 8-candidate tiles)
        if (short_last) {
            if (io.geom_o.nchunks > 1) return launch5<8, 1, 2, 4096 + 16384>(st, kind, d_gp, d_prog, io, grid, D);
            return launch5<8, 1, 2, 16384>(st, kind, d_gp, d_prog, io, grid, D);
        }
        if (io.geom_o.nchunks > 1) return launch5<8, 1, 2, 4096>(st, kind, d_gp, d_prog, io, grid, D);
        return launch5<8, 1, 2>(st, kind, d_gp, d_prog, io, grid, D);
    }
#ifdef GPFO_BUILD_EXPERIMENTS
    if (variant == 29) return launch5<8, 2, 2, 0, 8>(st, kind, d_gp, d_prog, io, grid, D);     // 8 warps, 16 candidates x 512-row chunks, two CTAs per SM                 // ablations of variant 20 (timing only; 22: no generation, 23: no L^-1 copies,
    if (variant == 22) return launch5<8, 2, 2, 1>(st, kind, d_gp, d_prog, io, grid, D);      // 24: neither, 25: no mbarrier waits,
    if (variant == 23) return launch5<8, 2, 2, 2>(st, kind, d_gp, d_prog, io, grid, D);      // 26: none of the three)
    if (variant == 24) return launch5<8, 2, 2, 3>(st, kind, d_gp, d_prog, io, grid, D);
    if (variant == 25) return launch5<8, 2, 2, 4>(st, kind, d_gp, d_prog, io, grid, D);
    if (variant == 26) return launch5<8, 2, 2, 7>(st, kind, d_gp, d_prog, io, grid, D);
    if (variant == 27) return launch5<8, 2, 2, 8>(st, kind, d_gp, d_prog, io, grid, D);      // 27 / 28: phase cycle counters of 20 / 21
    if (variant == 28) return launch5<4, 4, 2, 8>(st, kind, d_gp, d_prog, io, grid, D);
    if (variant == 30) return launch5<8, 2, 2, 16>(st, kind, d_gp, d_prog, io, grid, D);     // scheduling experiments: every warp
    if (variant == 31) return launch5<4, 4, 2, 16>(st, kind, d_gp, d_prog, io, grid, D);     // generates first (30 / 31), one group
    if (variant == 32) return launch5<8, 2, 2, 32>(st, kind, d_gp, d_prog, io, grid, D);     // of lookahead (32 / 33), every warp
    if (variant == 33) return launch5<4, 4, 2, 32>(st, kind, d_gp, d_prog, io, grid, D);     // generates last (34 / 35)
    if (variant == 34) return launch5<8, 2, 2, 64>(st, kind, d_gp, d_prog, io, grid, D);
    if (variant == 35) return launch5<4, 4, 2, 64>(st, kind, d_gp, d_prog, io, grid, D);
    if (variant == 36) {                                    // 36 / 37: 20 / 21 with one group per contraction step (before the wide steps)
        if (io.geom_o.nchunks > 1) return launch5<8, 2, 2, 2048 + 4096>(st, kind, d_gp, d_prog, io, grid, D);
        return launch5<8, 2, 2, 2048>(st, kind, d_gp, d_prog, io, grid, D);
    }
    if (variant == 37) return launch5<4, 4, 2, 2048>(st, kind, d_gp, d_prog, io, grid, D);
    if (variant == 38) return launch5<4, 4, 2, 1>(st, kind, d_gp, d_prog, io, grid, D);      // 38 / 39: ablations of 21 (no generation / no L^-1 copies)
    if (variant == 39) return launch5<4, 4, 2, 2>(st, kind, d_gp, d_prog, io, grid, D);
#endif
    // tiles of several chunks (M > 1024): row slots balanced over the schedulers (ABL bit 4096, contract5_kernel.cuh)
    if (short_last) {
        if (io.geom_o.nchunks > 1) return launch5<8, 2, 2, 4096 + 16384>(st, kind, d_gp, d_prog, io, grid, D);
        return launch5<8, 2, 2, 16384>(st, kind, d_gp, d_prog, io, grid, D);
    }
    if (io.geom_o.nchunks > 1) return launch5<8, 2, 2, 4096>(st, kind, d_gp, d_prog, io, grid, D);
    return launch5<8, 2, 2>(st, kind, d_gp, d_prog, io, grid, D);
}

// largest input dimension the shared-memory budget of a variant allows (two streaming buffers at least)
int gpfo_contract5_max_dim(int variant) {
    int D = GPFO_MAX_DIM;
    if (variant == 21) { while (D > 0 && Contract5Cfg<4, 4, 2>::smem_bytes(D, 2) > (size_t)227 * 1024) --D; }
    else { while (D > 0 && Contract5Cfg<8, 2, 2>::smem_bytes(D, 2) > (size_t)227 * 1024) --D; }
    return D;
}
