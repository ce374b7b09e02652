// Profiling builds of the contraction kernel (ablations, phase counters, scheduling experiments); never selected
// automatically.  41-43: the 32-candidate tiling without generation / without L^-1 loads / neither; 121-123: the same for
// the default 64-candidate tiling; 125: every warp generates before its DMMA work; 127: per-phase cycle counters.
#include "contract2_kernel.cuh"

cudaError_t gpfo_contract2_prof_launch(cudaStream_t st, int variant, int kind, const GpDev* d_gp, const ProgramDev* d_prog,
                                       EvalIO io, int grid, int D) {
    switch (variant) {
        case 121: return launch2<16, 2, 8, S12, 9>(st, kind, d_gp, d_prog, io, grid, D);
        case 122: return launch2<16, 2, 8, S12, 10>(st, kind, d_gp, d_prog, io, grid, D);
        case 123: return launch2<16, 2, 8, S12, 11>(st, kind, d_gp, d_prog, io, grid, D);
        case 127: return launch2<16, 2, 8, S12, 8 + 64>(st, kind, d_gp, d_prog, io, grid, D);
        case 125: return launch2<16, 2, 8, S12, 8 + 256>(st, kind, d_gp, d_prog, io, grid, D);
        case 41: return launch2<16, 4, 4, 4, 1>(st, kind, d_gp, d_prog, io, grid, D);
        case 42: return launch2<16, 4, 4, 4, 2>(st, kind, d_gp, d_prog, io, grid, D);
        case 43: return launch2<16, 4, 4, 4, 3>(st, kind, d_gp, d_prog, io, grid, D);
        default: return cudaErrorInvalidValue;
    }
}
