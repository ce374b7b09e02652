// Contraction kernel v5, launches of the stored-A gradient pass: the forward launch that parks A = L^-1 Kx per tile and the
// tall-chunk backward launch.  (Its own translation unit: the instantiations compile next to those of contract5.cu.)
#include "contract5_kernel.cuh"
#include "contract5_bwd_kernel.cuh"

// forward launch of the stored-A gradient pass (large batches): variant 20's tiling, A parked in the layout the backward launch
// (contract2_kernel, 64-candidate tiles) reads; the grid walks 16-candidate tiles
// (`native` = 1: A in the 16-candidate tile layout the tall-chunk backward kernel reads, one contiguous 8 KB block per group)
// (`variant` 21: the 32-candidate x 512-row tiling for small training sets, A always in the 64-candidate tile layout)
cudaError_t gpfo_contract5_astore_launch(cudaStream_t st, int kind, const GpDev* d_gp, const ProgramDev* d_prog, EvalIO io,
                                         int num_sms, int D, int native, int variant) {
    const bool short_last = io.geom_o.nrb % io.geom_o.rbc != 0;      // as gpfo_contract5_launch
    if (variant == 21) {
        const int64_t nt = (io.N + 31) / 32;
        const int g21 = (int)(nt < num_sms ? (nt > 0 ? nt : 1) : num_sms);
        if (short_last) return launch5<4, 4, 2, 128 + 16384>(st, kind, d_gp, d_prog, io, g21, D);
        return launch5<4, 4, 2, 128>(st, kind, d_gp, d_prog, io, g21, D);
    }
    const int64_t ntiles = (io.N + 15) / 16;
    const int grid = (int)(ntiles < num_sms ? (ntiles > 0 ? ntiles : 1) : num_sms);
    if (native) {
        if (short_last) {
            if (io.geom_o.nchunks > 1) return launch5<8, 2, 2, 128 + 512 + 4096 + 16384>(st, kind, d_gp, d_prog, io, grid, D);
            return launch5<8, 2, 2, 128 + 512 + 16384>(st, kind, d_gp, d_prog, io, grid, D);
        }
        if (io.geom_o.nchunks > 1) return launch5<8, 2, 2, 128 + 512 + 4096>(st, kind, d_gp, d_prog, io, grid, D);
        return launch5<8, 2, 2, 128 + 512>(st, kind, d_gp, d_prog, io, grid, D);
    }
    if (short_last) return launch5<8, 2, 2, 128 + 16384>(st, kind, d_gp, d_prog, io, grid, D);
    return launch5<8, 2, 2, 128>(st, kind, d_gp, d_prog, io, grid, D);
}

// backward launch of the stored-A pass in the tall-chunk tiling (A in the 16-candidate tile layout, `native` forward above)
cudaError_t gpfo_contract5_bwd_launch(cudaStream_t st, int kind, const GpDev* d_gp, EvalIO io, int num_sms, int D) {
    const int64_t ntiles = (io.N + 15) / 16;
    const int grid = (int)(ntiles < num_sms ? (ntiles > 0 ? ntiles : 1) : num_sms);
    switch (kind) {
        case GPFO_KERN_RBF: return launch5_bwd_kind<8, 2, 2, GPFO_KERN_RBF>(st, d_gp, io, grid, D);
        case GPFO_KERN_MATERN52: return launch5_bwd_kind<8, 2, 2, GPFO_KERN_MATERN52>(st, d_gp, io, grid, D);
        default: return launch5_bwd_kind<8, 2, 2, GPFO_KERN_MATERN32>(st, d_gp, io, grid, D);
    }
}
// does the tall-chunk backward kernel take inputs of dimension D?  (shared memory; partial tiles in the drained ring)
int gpfo_contract5_bwd_ok(int D) {
    return Contract5BwdCfg<8, 2, 2>::fits(D) && Contract5BwdCfg<8, 2, 2>::smem_bytes(D, 2) <= (size_t)227 * 1024;
}
// doubles of A per 16-candidate tile in the layout the two launches above share: 8 KB per group of 64 stream rows
long long gpfo_contract5_astore_stride(int nrb) { return (long long)((nrb + 7) / 8) * Contract5Cfg<8, 2, 2>::GROUP_SLOTS; }

