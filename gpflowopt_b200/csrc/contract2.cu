// Contraction kernel v2, forward launches of the production variants (kernel: contract2_kernel.cuh).
#include "contract2_kernel.cuh"

#ifdef GPFO_BUILD_EXPERIMENTS
cudaError_t gpfo_contract2_prof_launch(cudaStream_t st, int variant, int kind, const GpDev* d_gp, const ProgramDev* d_prog,
                                       EvalIO io, int grid, int D);
#endif

// variants: see the table in contract.cu
cudaError_t gpfo_contract2_launch(cudaStream_t st, int variant, int kind, const GpDev* d_gp, const ProgramDev* d_prog,
                                  EvalIO io, int grid, int D) {
    if ((variant >= 41 && variant <= 43) || (variant >= 120 && variant <= 129))
#ifdef GPFO_BUILD_EXPERIMENTS
        return gpfo_contract2_prof_launch(st, variant, kind, d_gp, d_prog, io, grid, D);
#else
        return cudaErrorInvalidValue;
#endif
    switch (variant) {
        case 5: return launch2<16, 4, 2, 4>(st, kind, d_gp, d_prog, io, grid, D);
        case 11: return launch2<16, 4, 4, 4, 8>(st, kind, d_gp, d_prog, io, grid, D);
        case 12: return launch2<16, 2, 8, S12, 8>(st, kind, d_gp, d_prog, io, grid, D);
#ifdef GPFO_BUILD_EXPERIMENTS
        case 14: return launch2<16, 2, 8, S12, 8 + 1024>(st, kind, d_gp, d_prog, io, grid, D);
#endif
        case 6: return launch2<16, 4, 1, 4>(st, kind, d_gp, d_prog, io, grid, D);
        default: return launch2<16, 4, 4, 4>(st, kind, d_gp, d_prog, io, grid, D);
    }
}
