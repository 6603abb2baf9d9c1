// Host-callable launchers of the sm_100a reshape kernels (implemented in kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "blocks.h"

namespace dtfftb {

// Tile configuration of the family-T (shared-memory transpose) kernel.
struct TileCfg {
    int ka;    // tile extent along a = 32*ka elements
    int kb;    // tile extent along b = 32*kb elements
    int rows;  // threadIdx.y extent (threads = 32*rows)
};

// Family T: out[b-contiguous] <- in[a-contiguous] for every block of the table.
// `es` = element bytes (4, 8, 16).  Table lives in device memory.
// `noshift`: ignore BlockDesc::bshift (a base pointer of this launch is not 128-byte aligned, so the shifted tile grid
// would not align anything; the table still covers every element, a few tiles are empty).
cudaError_t launch_transpose(int es, TileCfg cfg, const void* in, void* out, const BlockDesc* d_blocks,
                             int nblocks, long long total_items, int grid_cap, cudaStream_t stream, bool noshift = false);

// Family R: row copy in units of `unit` bytes (4, 8, 16); `tx` = threads along the row
// (power of two, 8..256).  Descriptors are pre-scaled to units.
cudaError_t launch_rows(int unit, int tx, const void* in, void* out, const BlockDesc* d_blocks, int nblocks,
                        long long total_items, int grid_cap, cudaStream_t stream);

// Rows each thread moves per tile in family R (tile = tx units x (256/tx)*kRowsPerThread rows).
constexpr int kRowsPerThread = 8;
constexpr int kRowsThreads = 256;

bool transpose_cfg_supported(int es, TileCfg cfg);

}  // namespace dtfftb
