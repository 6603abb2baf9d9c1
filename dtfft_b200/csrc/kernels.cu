// sm_100a reshape kernels: family T (tiled transpose) and family R (row copy).
//
// These replace the reference's run-time generated template
// (src/dtfft_nvrtc_module.F90:434-583: one CTA per (32x32 tile, plane), one element per
// thread per access, one launch PER PEER for pack/unpack).  Here:
//   * one launch covers every peer: CTAs walk a flattened work-item space over a table
//     of BlockDesc (blocks.h) with a grid-stride (persistent) loop;
//   * all global accesses of family R are 128-bit whenever the geometry allows
//     (the builder picks the widest unit that divides every stride/offset);
//   * family T stages a tile through padded shared memory: lanes run along the
//     input-contiguous axis on loads and along the output-contiguous axis on stores, so
//     both sides are fully coalesced; the +1 element row padding makes the transposed
//     shared-memory writes conflict-free for 4-, 8- and 16-byte elements;
//   * every load of a tile is issued before the first shared-memory store (registers
//     as staging) so each thread keeps several independent 16-byte requests in flight.
// No arithmetic touches the payload: results are bit-exact by construction.
#include "kernels.cuh"

#include <cstdlib>

namespace dtfftb {

namespace {

constexpr int kMaxDevices = 64;

template <int ES> struct ElemT;
template <> struct ElemT<4> { using type = unsigned int; };
template <> struct ElemT<8> { using type = uint2; };
template <> struct ElemT<16> { using type = uint4; };

// Payload accesses are plain C++ loads / stores under their bounds test (the compiler turns them into predicated
// LDG / STG and batches them freely); block bases come out of the descriptor table (local or peer-mapped HBM), so the
// kernels tell the compiler that they are global-space addresses -- __builtin_assume(__isGlobal(p)) -- which keeps them
// LDG / STG instead of generic LD / ST.  (Inline-asm ld.global / st.global was measured too: a C++ `if` around the asm put
// a branch around every load (4-byte permutes -7 %), the predicate inside the asm cost the 8-byte permutes 15 % but is
// the fastest form for 4-byte elements (6.6 vs 6.1-6.4 TB/s: 16 loads per thread); profiles/r02k_kbench_quick.txt,
// r02_kbench_quick.txt, r02l_kbench_quick_isglobal.txt.  Hence: 4-byte tiles use the predicated-asm accessors below,
// 8- and 16-byte tiles plain C++.)
__device__ __forceinline__ unsigned ld_global_if(const unsigned* p, bool pred) {
    unsigned v;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.global.u32 %0, [%1];\n\t}" : "=r"(v) : "l"(p), "r"(pred ? 1u : 0u));
    return v;
}
__device__ __forceinline__ void st_global_if(unsigned* p, unsigned v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u32 [%0], %1;\n\t}" ::"l"(p), "r"(v), "r"(pred ? 1u : 0u) : "memory");
}

__device__ __forceinline__ unsigned fast_div(unsigned n, const FastDiv& f) {
    return f.mul ? (__umulhi(n, f.mul) >> f.shr) : n;
}

// Last block whose item_begin <= item.
__device__ __forceinline__ int find_block(const BlockDesc* __restrict__ blocks, int nblocks, long long item) {
    int lo = 0, hi = nblocks - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(&blocks[mid].item_begin) <= item)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

struct ItemPos {
    int t0, t1, c;
};

__device__ __forceinline__ ItemPos decode_item(const BlockDesc& d, long long item) {
    unsigned local = (unsigned)(item - d.item_begin);
    unsigned q0 = fast_div(local, d.div0);
    ItemPos p;
    p.t0 = (int)(local - q0 * (unsigned)d.tiles0);
    unsigned q1 = fast_div(q0, d.div1);
    p.t1 = (int)(q0 - q1 * (unsigned)d.tiles1);
    p.c = (int)q1;
    return p;
}

// Fused NVLink tables carry the address of the sticky peer-error word (blocks.h: BlockDesc::abort).  Once
// a device barrier has timed out the destination was never announced free (or blocks never landed):
// every later kernel of the group stores nothing, and the next API call returns
// DTFFTB_ERROR_PEER_TIMEOUT.  One read per CTA, shared so that the whole CTA takes the same branch.
__device__ __forceinline__ bool peer_abort(const BlockDesc* __restrict__ blocks, int t) {
    const unsigned long long* flag = blocks[0].abort;
    if (!flag) return false;
    __shared__ unsigned s_abort;
    if (t == 0) s_abort = *reinterpret_cast<const volatile unsigned long long*>(flag) != 0ull ? 1u : 0u;
    __syncthreads();
    return s_abort != 0u;
}

// ---------------------------------------------------------------------------------
// Family T
// ---------------------------------------------------------------------------------
template <typename T, int KA, int KB, int ROWS>
__global__ void __launch_bounds__(32 * ROWS)
    transpose_tiles_kernel(const T* __restrict__ in, T* __restrict__ out, const BlockDesc* __restrict__ blocks,
                           int nblocks, long long total_items, int noshift) {
    constexpr int TA = 32 * KA, TB = 32 * KB;
    constexpr int PITCH = TB + 1;
    constexpr int LD_ROWS = TB / ROWS;  // b-rows each thread loads
    constexpr int ST_ROWS = TA / ROWS;  // a-rows each thread stores
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);  // tile[a][b], pitch PITCH

    const int tx = threadIdx.x, ty = threadIdx.y;
    const bool aborted = peer_abort(blocks, ty * 32 + tx);

    // Fused NVLink tables interleave the peers: consecutive CTAs take items spread over the whole
    // item space, so remote (NVLink-bound) and local (HBM-bound) tiles overlap instead of running
    // one peer after the other, and no peer sees the traffic of every rank at once.
    const long long shuffle = __ldg(&blocks[0].shuffle);
    for (long long it0 = blockIdx.x; it0 < total_items && !aborted; it0 += gridDim.x) {
        const long long item = shuffle > 1 ? (it0 * shuffle) % total_items : it0;
        const int bi = find_block(blocks, nblocks, item);
        const BlockDesc& d = blocks[bi];
        const ItemPos p = decode_item(d, item);
        const int n0 = d.n0, n1 = d.n1;
        const long long is1 = d.is1, os0 = d.os0;
        const T* src = (d.in_base ? reinterpret_cast<const T*>(d.in_base) : in) + d.in_off + (long long)p.c * d.is2;
        T* dst = (d.out_base ? reinterpret_cast<T*>(d.out_base) : out) + d.out_off + (long long)p.c * d.os2;
        __builtin_assume(__isGlobal(src));
        __builtin_assume(__isGlobal(dst));
        // Destination runs that start off a 128-byte line (uneven splits: 250 x 8 B ...) would make every tile boundary
        // split a line between two warps: partial-line stores, ruinous over NVLink (config 5, Y->Z: 204 GB/s).  The
        // table builder shifts the tile grid back by d.bshift elements instead; the first tile is masked at its start.
        const int a0 = p.t0 * TA, b0 = p.t1 * TB - (noshift ? 0 : d.bshift);

        T regs[LD_ROWS][KA];
#pragma unroll
        for (int j = 0; j < LD_ROWS; ++j) {
            const int b = b0 + j * ROWS + ty;
#pragma unroll
            for (int k = 0; k < KA; ++k) {
                const int a = a0 + tx + 32 * k;
                if constexpr (sizeof(T) == 4) {
                    regs[j][k] = ld_global_if(src + a + (long long)b * is1, a < n0 && b >= 0 && b < n1);
                } else {
                    if (a < n0 && b >= 0 && b < n1) regs[j][k] = src[a + (long long)b * is1];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < LD_ROWS; ++j)
#pragma unroll
            for (int k = 0; k < KA; ++k) tile[(tx + 32 * k) * PITCH + j * ROWS + ty] = regs[j][k];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ST_ROWS; ++j) {
            const int a = a0 + j * ROWS + ty;
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int b = b0 + tx + 32 * k;
                if constexpr (sizeof(T) == 4) {
                    st_global_if(dst + (long long)a * os0 + b, tile[(j * ROWS + ty) * PITCH + tx + 32 * k], a < n0 && b >= 0 && b < n1);
                } else {
                    if (a < n0 && b >= 0 && b < n1) dst[(long long)a * os0 + b] = tile[(j * ROWS + ty) * PITCH + tx + 32 * k];
                }
            }
        }
        __syncthreads();
    }
}

template <int ES, int KA, int KB, int ROWS>
cudaError_t launch_T(const void* in, void* out, const BlockDesc* blocks, int nblocks, long long total, int grid_cap,
                     cudaStream_t stream, int noshift) {
    using T = typename ElemT<ES>::type;
    constexpr size_t smem = (size_t)(32 * KA) * (32 * KB + 1) * ES;
    auto kern = transpose_tiles_kernel<T, KA, KB, ROWS>;
    if (smem > 48 * 1024) {
        // function attributes are per device: one flag per (instantiation, device)
        static bool attr_set[kMaxDevices] = {};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            // let three 66.5 KB tiles share an SM
            e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
        }
    }
    long long g = total < grid_cap ? total : grid_cap;
    kern<<<(unsigned)g, dim3(32, ROWS), smem, stream>>>(reinterpret_cast<const T*>(in), reinterpret_cast<T*>(out),
                                                        blocks, nblocks, total, noshift);
    return cudaGetLastError();
}

template <int ES>
cudaError_t dispatch_T(TileCfg c, const void* in, void* out, const BlockDesc* b, int nb, long long total, int cap,
                       cudaStream_t s, int noshift) {
#define DTFFTB_T_CASE(KA_, KB_, R_) \
    if (c.ka == KA_ && c.kb == KB_ && c.rows == R_) return launch_T<ES, KA_, KB_, R_>(in, out, b, nb, total, cap, s, noshift);
    DTFFTB_T_CASE(1, 1, 4)
    DTFFTB_T_CASE(1, 1, 8)
    DTFFTB_T_CASE(1, 1, 16)
    DTFFTB_T_CASE(2, 1, 8)
    DTFFTB_T_CASE(1, 2, 8)
    DTFFTB_T_CASE(2, 2, 8)
    DTFFTB_T_CASE(2, 2, 16)
    DTFFTB_T_CASE(1, 4, 16)  // 32 x 128: 2 KB contiguous store runs at 16 B / element (round-2 sweep candidates)
    DTFFTB_T_CASE(4, 1, 16)  // 128 x 32: 2 KB contiguous load runs
#undef DTFFTB_T_CASE
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------
// Family R
// ---------------------------------------------------------------------------------
template <typename V, int TX>
__global__ void __launch_bounds__(kRowsThreads)
    rows_copy_kernel(const V* __restrict__ in, V* __restrict__ out, const BlockDesc* __restrict__ blocks, int nblocks,
                     long long total_items) {
    constexpr int TY = kRowsThreads / TX;
    constexpr int UR = kRowsPerThread;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const bool aborted = peer_abort(blocks, (int)threadIdx.x);

    // Fused NVLink tables interleave the peers: consecutive CTAs take items spread over the whole
    // item space, so remote (NVLink-bound) and local (HBM-bound) tiles overlap instead of running
    // one peer after the other, and no peer sees the traffic of every rank at once.
    const long long shuffle = __ldg(&blocks[0].shuffle);
    for (long long it0 = blockIdx.x; it0 < total_items && !aborted; it0 += gridDim.x) {
        const long long item = shuffle > 1 ? (it0 * shuffle) % total_items : it0;
        const int bi = find_block(blocks, nblocks, item);
        const BlockDesc& d = blocks[bi];
        const ItemPos p = decode_item(d, item);
        const int n0 = d.n0, n1 = d.n1;
        const long long is1 = d.is1, os1 = d.os1;
        const V* src = (d.in_base ? reinterpret_cast<const V*>(d.in_base) : in) + d.in_off + (long long)p.c * d.is2;
        V* dst = (d.out_base ? reinterpret_cast<V*>(d.out_base) : out) + d.out_off + (long long)p.c * d.os2;
        __builtin_assume(__isGlobal(src));
        __builtin_assume(__isGlobal(dst));
        const int col = p.t0 * TX + tx;
        const int row0 = p.t1 * (TY * UR) + ty;
        if (col < n0) {
            V regs[UR];
#pragma unroll
            for (int r = 0; r < UR; ++r) {
                const int row = row0 + r * TY;
                if (row < n1) regs[r] = src[col + (long long)row * is1];
            }
#pragma unroll
            for (int r = 0; r < UR; ++r) {
                const int row = row0 + r * TY;
                if (row < n1) dst[col + (long long)row * os1] = regs[r];
            }
        }
    }
}

template <int U, int TX>
cudaError_t launch_R(const void* in, void* out, const BlockDesc* blocks, int nblocks, long long total, int grid_cap,
                     cudaStream_t stream) {
    using V = typename ElemT<U>::type;
    long long g = total < grid_cap ? total : grid_cap;
    const V* pin = reinterpret_cast<const V*>(in);
    V* pout = reinterpret_cast<V*>(out);
    rows_copy_kernel<V, TX><<<(unsigned)g, kRowsThreads, 0, stream>>>(pin, pout, blocks, nblocks, total);
    return cudaGetLastError();
}

template <int U>
cudaError_t dispatch_R(int tx, const void* in, void* out, const BlockDesc* b, int nb, long long total, int cap,
                       cudaStream_t s) {
    switch (tx) {
        case 8: return launch_R<U, 8>(in, out, b, nb, total, cap, s);
        case 16: return launch_R<U, 16>(in, out, b, nb, total, cap, s);
        case 32: return launch_R<U, 32>(in, out, b, nb, total, cap, s);
        case 64: return launch_R<U, 64>(in, out, b, nb, total, cap, s);
        case 128: return launch_R<U, 128>(in, out, b, nb, total, cap, s);
        case 256: return launch_R<U, 256>(in, out, b, nb, total, cap, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

bool transpose_cfg_supported(int es, TileCfg c) {
    if (es != 4 && es != 8 && es != 16) return false;
    const int ok[][3] = {{1, 1, 4}, {1, 1, 8}, {1, 1, 16}, {2, 1, 8}, {1, 2, 8}, {2, 2, 8}, {2, 2, 16}, {1, 4, 16}, {4, 1, 16}};
    for (auto& o : ok)
        if (c.ka == o[0] && c.kb == o[1] && c.rows == o[2]) return true;
    return false;
}

cudaError_t launch_transpose(int es, TileCfg cfg, const void* in, void* out, const BlockDesc* d_blocks, int nblocks,
                             long long total_items, int grid_cap, cudaStream_t stream, bool noshift) {
    if (total_items <= 0) return cudaSuccess;
    switch (es) {
        case 4: return dispatch_T<4>(cfg, in, out, d_blocks, nblocks, total_items, grid_cap, stream, noshift ? 1 : 0);
        case 8: return dispatch_T<8>(cfg, in, out, d_blocks, nblocks, total_items, grid_cap, stream, noshift ? 1 : 0);
        case 16: return dispatch_T<16>(cfg, in, out, d_blocks, nblocks, total_items, grid_cap, stream, noshift ? 1 : 0);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_rows(int unit, int tx, const void* in, void* out, const BlockDesc* d_blocks, int nblocks,
                        long long total_items, int grid_cap, cudaStream_t stream) {
    if (total_items <= 0) return cudaSuccess;
    switch (unit) {
        case 4: return dispatch_R<4>(tx, in, out, d_blocks, nblocks, total_items, grid_cap, stream);
        case 8: return dispatch_R<8>(tx, in, out, d_blocks, nblocks, total_items, grid_cap, stream);
        case 16: return dispatch_R<16>(tx, in, out, d_blocks, nblocks, total_items, grid_cap, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace dtfftb
