// Host-side kernel object (see kernel_object.h).  Geometry of every kernel kind follows the
// reference's index maps (src/dtfft_nvrtc_module.F90:494-578 for the device strings,
// src/include/_dtfft_kernel_host_routines.inc for the host loops); each case below cites
// the lines it restates.
#include "kernel_object.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/dtfft_b200.h"
#include "errors.h"

namespace dtfftb {

bool is_per_neighbor_kind(int t) {
    switch (t) {
        case K_UNPACK_PIPELINED:
        case K_PERMUTE_BACKWARD_END_PIPELINED:
        case K_COPY_PIPELINED:
        case K_PACK_BACKWARD:
        case K_PACK_FORWARD:
        case K_PACK_PIPELINED:
        case K_UNPACK_FORWARD_PIPELINED:
        case K_UNPACK_BACKWARD_PIPELINED: return true;
        default: return false;
    }
}

bool needs_neighbor_data(int t) {  // is_pack_kernel || is_unpack_kernel, abstract_kernel.F90:100-104
    switch (t) {
        case K_PERMUTE_FORWARD:
        case K_PERMUTE_BACKWARD:
        case K_PERMUTE_BACKWARD_START:
        case K_COPY:
        case K_DUMMY: return false;
        default: return true;
    }
}

Family family_of(int t) {
    switch (t) {
        case K_COPY:
        case K_COPY_PIPELINED: return FAM_COPY;
        case K_PERMUTE_FORWARD:
        case K_PERMUTE_BACKWARD:
        case K_PERMUTE_BACKWARD_START:
        case K_PACK_FORWARD:
        case K_PACK_BACKWARD:
        case K_UNPACK_FORWARD:
        case K_UNPACK_FORWARD_PIPELINED:
        case K_UNPACK_BACKWARD:
        case K_UNPACK_BACKWARD_PIPELINED: return FAM_T;
        case K_PACK:
        case K_PACK_PIPELINED:
        case K_UNPACK:
        case K_UNPACK_PIPELINED:
        case K_PERMUTE_BACKWARD_END:
        case K_PERMUTE_BACKWARD_END_PIPELINED: return FAM_R;
        default: return FAM_NONE;
    }
}

int effective_type(int t, int ndims) {  // abstract_kernel.F90:271-283
    if (ndims != 2) return t;
    switch (t) {
        case K_PACK_BACKWARD: return K_PACK_FORWARD;
        case K_PERMUTE_BACKWARD: return K_PERMUTE_FORWARD;
        case K_UNPACK_BACKWARD: return K_UNPACK_FORWARD;
        case K_UNPACK_BACKWARD_PIPELINED: return K_UNPACK_FORWARD_PIPELINED;
        default: return t;
    }
}

Box make_box(int t, int ndims, const int32_t* dims, const int32_t* nd5) {
    const long long nx = dims[0], ny = dims[1], nz = ndims == 3 ? dims[2] : 1;
    long long nxx = 0, nyy = 0, nzz = 1, din = 0, dout = 0;
    if (nd5) {
        nxx = nd5[0];
        nyy = nd5[1];
        nzz = ndims == 3 ? nd5[2] : 1;
        din = nd5[3];
        dout = nd5[4];
    }
    Box b;
    switch (t) {
        case K_PERMUTE_FORWARD:  // nvrtc_module.F90:495-497,533,561 ; host .inc:140-196
            // out[y + z*ny + x*ny*nz] = in[x + y*nx + z*nx*ny]   (a=x, b=y, c=z)
            b.n0 = nx, b.n1 = ny, b.n2 = nz;
            b.is1 = nx, b.is2 = nx * ny;
            b.os0 = ny * nz, b.os1 = 1, b.os2 = ny;
            break;
        case K_PERMUTE_BACKWARD:  // :501-503,539,571 ; .inc:262-302
            // out[z + x*nz + y*nz*nx] = in[x + y*nx + z*nx*ny]   (a=x, b=z, c=y)
            b.n0 = nx, b.n1 = nz, b.n2 = ny;
            b.is1 = nx * ny, b.is2 = nx;
            b.os0 = nz, b.os1 = 1, b.os2 = nz * nx;
            break;
        case K_PERMUTE_BACKWARD_START:  // :507-509 ; .inc:351-390
            // out[z + y*nz + x*nz*ny] = in[x + y*nx + z*nx*ny]   (a=x, b=z, c=y)
            b.n0 = nx, b.n1 = nz, b.n2 = ny;
            b.is1 = nx * ny, b.is2 = nx;
            b.os0 = nz * ny, b.os1 = 1, b.os2 = nz;
            break;
        case K_PACK_FORWARD:  // :498-500,563 ; .inc:1026-1086
            // out[dout + y + z*nyy + x*nyy*nzz] = in[din + x + y*nx + z*nx*ny]
            b.n0 = nxx, b.n1 = nyy, b.n2 = nzz;
            b.is1 = nx, b.is2 = nx * ny;
            b.os0 = nyy * nzz, b.os1 = 1, b.os2 = nyy;
            b.in_off = din, b.out_off = dout;
            break;
        case K_PACK_BACKWARD:  // :504-506,569 ; .inc:1155-1198
            // out[dout + z + x*nzz + y*nzz*nxx] = in[din + x + y*nx + z*nx*ny]  (a=x, b=z, c=y)
            b.n0 = nxx, b.n1 = nzz, b.n2 = nyy;
            b.is1 = nx * ny, b.is2 = nx;
            b.os0 = nzz, b.os1 = 1, b.os2 = nzz * nxx;
            b.in_off = din, b.out_off = dout;
            break;
        case K_UNPACK_FORWARD_PIPELINED:  // host only in the reference: .inc:757-811
            if (ndims == 2) {
                // out[dout + x + y*nx] = in[din + y + x*nyy]             (a=y, b=x)
                b.n0 = nyy, b.n1 = nxx, b.n2 = 1;
                b.is1 = nyy, b.is2 = 0;
                b.os0 = nx, b.os1 = 1, b.os2 = 0;
            } else {
                // out[dout + x + y*nx + z*nx*ny] = in[din + z + x*nzz + y*nzz*nxx]  (a=z, b=x, c=y)
                b.n0 = nzz, b.n1 = nxx, b.n2 = nyy;
                b.is1 = nzz, b.is2 = nzz * nxx;
                b.os0 = nx * ny, b.os1 = 1, b.os2 = nx;
            }
            b.in_off = din, b.out_off = dout;
            break;
        case K_UNPACK_BACKWARD_PIPELINED:  // host only in the reference: .inc:907-946
            // out[dout + x + y*nx + z*nx*ny] = in[din + y + z*nyy + x*nzz*nyy]      (a=y, b=x, c=z)
            b.n0 = nyy, b.n1 = nxx, b.n2 = nzz;
            b.is1 = nzz * nyy, b.is2 = nyy;
            b.os0 = nx, b.os1 = 1, b.os2 = nx * ny;
            b.in_off = din, b.out_off = dout;
            break;
        case K_UNPACK_PIPELINED:  // :513-515,537,565 ; .inc:575-636
            // out[dout + x + y*nx + z*nx*ny] = in[din + x + y*nxx + z*nxx*nyy]
            b.n0 = nxx, b.n1 = nyy, b.n2 = nzz;
            b.is1 = nxx, b.is2 = nxx * nyy;
            b.os0 = 1, b.os1 = nx, b.os2 = nx * ny;
            b.in_off = din, b.out_off = dout;
            break;
        case K_PACK_PIPELINED:  // :516-518,533,567 ; .inc:657-721
            // out[dout + x + y*nxx + z*nxx*nyy] = in[din + x + y*nx + z*nx*ny]
            b.n0 = nxx, b.n1 = nyy, b.n2 = nzz;
            b.is1 = nx, b.is2 = nx * ny;
            b.os0 = 1, b.os1 = nxx, b.os2 = nxx * nyy;
            b.in_off = din, b.out_off = dout;
            break;
        case K_PERMUTE_BACKWARD_END_PIPELINED:  // :510-512,535,565 ; .inc:439-482
            // out[dout + x + y*nx + z*nx*ny] = in[din + x + z*nxx + y*nxx*nzz]
            b.n0 = nxx, b.n1 = nyy, b.n2 = nzz;
            b.is1 = nxx * nzz, b.is2 = nxx;
            b.os0 = 1, b.os1 = nx, b.os2 = nx * ny;
            b.in_off = din, b.out_off = dout;
            break;
        default: break;
    }
    return b;
}

namespace {

int per_neighbor_type(int t) {
    switch (t) {
        case K_PACK: return K_PACK_PIPELINED;
        case K_UNPACK: return K_UNPACK_PIPELINED;
        case K_PERMUTE_BACKWARD_END: return K_PERMUTE_BACKWARD_END_PIPELINED;
        case K_UNPACK_FORWARD: return K_UNPACK_FORWARD_PIPELINED;
        case K_UNPACK_BACKWARD: return K_UNPACK_BACKWARD_PIPELINED;
        default: return t;
    }
}

// B200 tile table, measured at 512^3 with tools/kbench.py (profiles/r01b_kbench.md); replaces the
// Volta/Ampere model of nvrtc_block_optimizer.  16-byte elements want the smaller 32x64 tile
// (33 KB smem, 8 independent 16-B loads per thread, ~5 CTAs/SM): 6.8 TB/s vs 5.4 TB/s for the
// 64x64 tile whose 66.5 KB / 100 registers cap residency at 2 CTAs/SM.
TileCfg default_tile(int es) {
    switch (es) {
        case 16: return TileCfg{1, 2, 8};
        case 8: return TileCfg{2, 2, 16};
        default: return TileCfg{2, 2, 8};
    }
}

int pow2_ceil(long long v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Family R: drop unit axes / merge contiguous axes (in elements).
void normalize_rows(Box& b) {
    if (b.empty()) return;
    if (b.n2 > 1 && b.n1 > 1 && b.is2 == b.is1 * b.n1 && b.os2 == b.os1 * b.n1) {  // merge b,c
        b.n1 *= b.n2;
        b.n2 = 1;
        b.is2 = b.os2 = 0;
    }
    if (b.n1 == 1 && b.n2 > 1) {  // shift c into b
        b.n1 = b.n2, b.is1 = b.is2, b.os1 = b.os2;
        b.n2 = 1, b.is2 = b.os2 = 0;
    }
    if (b.n1 > 1 && b.is1 == b.n0 && b.os1 == b.n0) {  // rows contiguous on both sides: merge a,b
        b.n0 *= b.n1;
        b.n1 = b.n2, b.is1 = b.is2, b.os1 = b.os2;
        b.n2 = 1, b.is2 = b.os2 = 0;
        if (b.n1 > 1 && b.is1 == b.n0 && b.os1 == b.n0) {
            b.n0 *= b.n1;
            b.n1 = 1, b.is1 = b.os1 = 0;
        }
    }
}

int gcd_unit(long long bytes, int u) {
    while (u > 4 && (bytes % u) != 0) u >>= 1;
    return u;
}

}  // namespace

Kernel::~Kernel() { destroy(); }

void Kernel::destroy() {
    if (d_blocks_) cudaFree(d_blocks_);
    d_blocks_ = nullptr;
    created_ = false;
    noop_ = true;
    custom_ = false;
    nd_.clear();
    boxes_.clear();
    peer_out_.clear();
    peer_out_displ_.clear();
    for (int i = 0; i < 3; ++i) {
        all_[i] = DeviceTable{};
        single_[i].clear();
    }
}

int Kernel::create(int ndims, const int32_t* dims, int kernel_type, int64_t base_storage, const int32_t* neighbor_data,
                   int n_neighbors, int effort, bool force_effort) {
    (void)force_effort;
    destroy();
    if (ndims != 2 && ndims != 3) return DTFFT_ERROR_INVALID_N_DIMENSIONS;
    if (!dims) return DTFFT_ERROR_INVALID_USAGE;
    if (base_storage != 4 && base_storage != 8 && base_storage != 16) return DTFFTB_ERROR_INTERNAL;
    ndims_ = ndims;
    es_ = base_storage;
    for (int i = 0; i < 3; ++i) dims_[i] = i < ndims ? dims[i] : 1;
    created_ = true;
    // abstract_kernel.F90:236-245: zero-volume ranks and KERNEL_DUMMY are no-ops
    for (int i = 0; i < ndims; ++i)
        if (dims[i] < 0) return DTFFT_ERROR_INVALID_DIMENSION_SIZE;
    for (int i = 0; i < ndims; ++i)
        if (dims[i] == 0) {
            noop_ = true;
            type_ = kernel_type;
            return DTFFT_SUCCESS;
        }
    if (kernel_type == K_DUMMY) {
        noop_ = true;
        type_ = K_DUMMY;
        return DTFFT_SUCCESS;
    }
    type_ = effective_type(kernel_type, ndims);
    family_ = family_of(type_);
    if (family_ == FAM_NONE) return DTFFTB_ERROR_INTERNAL;
    if ((type_ == K_PERMUTE_BACKWARD_START || type_ == K_PERMUTE_BACKWARD_END ||
         type_ == K_PERMUTE_BACKWARD_END_PIPELINED || type_ == K_PERMUTE_BACKWARD) &&
        ndims != 3)
        return DTFFTB_ERROR_INTERNAL;  // abstract_kernel.F90:255-260 (debug check in the reference)
    if (needs_neighbor_data(type_)) {
        if (!neighbor_data || n_neighbors <= 0) return DTFFTB_ERROR_INTERNAL;  // "Neighbor data required"
        P_ = n_neighbors;
        nd_.assign(neighbor_data, neighbor_data + 5 * (size_t)n_neighbors);
    } else {
        P_ = 0;
    }
    noop_ = false;

    if (!dry_) {
        int dev = 0;
        cudaError_t ce = cudaGetDevice(&dev);
        if (ce != cudaSuccess) return cuda_error(ce);
        ce = cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, dev);
        if (ce != cudaSuccess) return cuda_error(ce);
    }

    if (family_ == FAM_COPY) return DTFFT_SUCCESS;

    const int ptype = per_neighbor_type(type_);
    if (P_ == 0) {
        boxes_.push_back(make_box(type_, ndims_, dims_, nullptr));
    } else {
        for (int n = 0; n < P_; ++n) boxes_.push_back(make_box(ptype, ndims_, dims_, &nd_[5 * (size_t)n]));
    }
    // int32 element-index limit of the reference (reshape_handle_generic.F90:159-172) is
    // lifted: descriptors are 64-bit.  Work items per block must still fit 31 bits.

    // B200 tile table (replaces nvrtc_block_optimizer's Volta/Ampere model).
    if (family_ == FAM_T) {
        tile_ = default_tile((int)es_);
        if (const char* e = getenv("DTFFTB_TILE")) {
            int ka, kb, r;
            if (sscanf(e, "%d,%d,%d", &ka, &kb, &r) == 3 && transpose_cfg_supported((int)es_, TileCfg{ka, kb, r}))
                tile_ = TileCfg{ka, kb, r};
        }
    }
    int rc = rebuild_tables();
    if (rc != DTFFT_SUCCESS) return rc;

    if (effort >= 3 /* DTFFT_EXHAUSTIVE */ && family_ == FAM_T && !dry_) {
        // Timed kernel autotune on scratch buffers (kernel_device.F90:338-397).
        long long elems = 0, in_need = 0, out_need = 0;
        for (auto& b : boxes_) {
            if (b.empty()) continue;
            elems += b.volume();
            in_need = std::max(in_need, b.in_off + (b.n0 - 1) + (b.n1 - 1) * b.is1 + (b.n2 - 1) * b.is2 + 1);
            out_need = std::max(out_need, b.out_off + (b.n0 - 1) * b.os0 + (b.n1 - 1) * b.os1 + (b.n2 - 1) * b.os2 + 1);
        }
        void *pi = nullptr, *po = nullptr;
        if (elems > 0 && cudaMalloc(&pi, in_need * es_) == cudaSuccess) {
            if (cudaMalloc(&po, out_need * es_) == cudaSuccess) {
                float ms = 0;
                autotune(pi, po, nullptr, 2, 5, &ms);
                cudaFree(po);
            }
            cudaFree(pi);
        }
        cudaGetLastError();
    }
    return DTFFT_SUCCESS;
}

int Kernel::create_boxes(Family family, int64_t base_storage, const std::vector<Box>& boxes) {
    destroy();
    if (family != FAM_T && family != FAM_R) return DTFFTB_ERROR_INTERNAL;
    if (base_storage != 4 && base_storage != 8 && base_storage != 16) return DTFFTB_ERROR_INTERNAL;
    es_ = base_storage;
    ndims_ = 3;
    created_ = true;
    custom_ = true;
    family_ = family;
    type_ = family == FAM_T ? K_PACK_FORWARD : K_PACK_PIPELINED;  // per-peer kinds: execute(n) = one peer, execute_all = every peer
    P_ = (int)boxes.size();
    boxes_ = boxes;
    bool any = false;
    for (auto& b : boxes_) any |= !b.empty();
    noop_ = !any;
    if (noop_) return DTFFT_SUCCESS;
    if (!dry_) {
        int dev = 0;
        cudaError_t ce = cudaGetDevice(&dev);
        if (ce != cudaSuccess) return cuda_error(ce);
        ce = cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, dev);
        if (ce != cudaSuccess) return cuda_error(ce);
    }
    if (family_ == FAM_T) {
        tile_ = default_tile((int)es_);
        if (const char* e = getenv("DTFFTB_TILE")) {
            int ka, kb, r;
            if (sscanf(e, "%d,%d,%d", &ka, &kb, &r) == 3 && transpose_cfg_supported((int)es_, TileCfg{ka, kb, r}))
                tile_ = TileCfg{ka, kb, r};
        }
    }
    return rebuild_tables();
}

int Kernel::rebuild_tables() {
    if (const char* e = getenv("DTFFTB_ALIGN_TILES")) align_b_ = atoi(e) != 0;
    if (d_blocks_) cudaFree(d_blocks_);
    d_blocks_ = nullptr;
    host_tables_.clear();
    std::vector<BlockDesc> host;
    // CTAs per resident slot.  Measured on B200 (profiles/r01_kbench.md): a static persistent
    // partition (1) loses ~7 % to SMs that finish early; one CTA per tile (0 = no cap, the
    // default) lets the hardware scheduler balance the tail and is the fastest setting.
    int grid_mult = 0;
    if (const char* e = getenv("DTFFTB_GRID_MULT")) grid_mult = std::max(0, atoi(e));

    auto make_desc = [&](const Box& b, int t0, int t1, long long begin, int peer) {
        BlockDesc d{};
        d.in_off = b.in_off, d.out_off = b.out_off;
        d.in_base = nullptr;
        d.out_base = (peer >= 0 && peer < (int)peer_out_.size()) ? peer_out_[peer] : nullptr;
        d.is1 = b.is1, d.is2 = b.is2, d.os0 = b.os0, d.os1 = b.os1, d.os2 = b.os2;
        d.n0 = (int)b.n0, d.n1 = (int)b.n1, d.n2 = (int)b.n2;
        d.tiles0 = (int)((b.n0 + t0 - 1) / t0);
        d.bshift = 0;
        if (family_ == FAM_T && align_b_) {
            // tiles along the output-contiguous axis start on 128-byte lines of the destination when every row of the box
            // has the same misalignment (outer strides are multiples of a line); bases are taken as line-aligned here and
            // checked at launch (Kernel::launch passes `noshift` otherwise)
            const long long L = 128 / es_;
            const bool rows_alike = (b.n0 == 1 || (b.os0 % L) == 0) && (b.n2 == 1 || (b.os2 % L) == 0);
            if (rows_alike && b.n1 * es_ >= 256) d.bshift = (int)(((b.out_off % L) + L) % L);
        }
        d.tiles1 = (int)((b.n1 + d.bshift + t1 - 1) / t1);
        d.div0 = FastDiv::make((unsigned)d.tiles0);
        d.div1 = FastDiv::make((unsigned)d.tiles1);
        d.item_begin = begin;
        return d;
    };
    auto items_of = [](const BlockDesc& d) { return (long long)d.tiles0 * d.tiles1 * d.n2; };

    if (family_ == FAM_T) {
        const int TA = 32 * tile_.ka, TB = 32 * tile_.kb;
        all_[0] = DeviceTable{};
        all_[0].offset = (long long)host.size();
        long long item = 0;
        for (size_t n = 0; n < boxes_.size(); ++n) {
            if (boxes_[n].empty()) continue;
            BlockDesc d = make_desc(boxes_[n], TA, TB, item, (int)n);
            if (items_of(d) >= (1ll << 31)) return DTFFTB_ERROR_INTERNAL;
            item += items_of(d);
            host.push_back(d);
            all_[0].nblocks++;
        }
        all_[0].total_items = item;
        single_[0].assign(boxes_.size(), DeviceTable{});
        for (size_t n = 0; n < boxes_.size(); ++n) {
            if (boxes_[n].empty()) continue;
            BlockDesc d = make_desc(boxes_[n], TA, TB, 0, (int)n);
            single_[0][n].offset = (long long)host.size();
            single_[0][n].nblocks = 1;
            single_[0][n].total_items = items_of(d);
            host.push_back(d);
        }
        const int threads = 32 * tile_.rows;
        const size_t smem = (size_t)TA * (TB + 1) * es_;
        int per_sm = std::min({2048 / threads, (int)((227 * 1024) / (smem + 1024)), 32});
        grid_cap_ = grid_mult ? sm_count_ * std::max(1, per_sm) * grid_mult : 0x7fffffff;
    } else if (family_ == FAM_R) {
        // widest unit allowed by the geometry
        std::vector<Box> norm = boxes_;
        long long max_row_bytes = 0;
        unit_geo_ = 16;
        for (auto& b : norm) {
            if (b.empty()) continue;
            normalize_rows(b);
            const long long q[] = {b.n0, b.is1, b.is2, b.os1, b.os2, b.in_off, b.out_off};
            for (long long v : q) unit_geo_ = gcd_unit(v * es_, unit_geo_);
            max_row_bytes = std::max(max_row_bytes, b.n0 * es_);
        }
        for (size_t n = 0; n < peer_out_displ_.size(); ++n) unit_geo_ = gcd_unit(peer_out_displ_[n] * es_, unit_geo_);
        int slot = 0;
        for (int unit = 4; unit <= 16; unit <<= 1, ++slot) {
            all_[slot] = DeviceTable{};
            single_[slot].assign(boxes_.size(), DeviceTable{});
            if (unit > unit_geo_) continue;
            const int tx = std::min(256, std::max(8, pow2_ceil(max_row_bytes / unit)));
            const int ty = kRowsThreads / tx;
            // scale boxes to units, re-row fully contiguous ones
            std::vector<Box> scaled;      // boxes of the combined table
            std::vector<int> owner;       // neighbour each scaled box belongs to
            for (size_t n = 0; n < norm.size(); ++n) {
                Box b = norm[n];
                if (b.empty()) continue;
                const long long f = es_;  // bytes per element
                b.n0 = b.n0 * f / unit;
                b.is1 = b.is1 * f / unit, b.is2 = b.is2 * f / unit;
                b.os1 = b.os1 * f / unit, b.os2 = b.os2 * f / unit;
                b.in_off = b.in_off * f / unit, b.out_off = b.out_off * f / unit;
                b.os0 = 1;
                const long long W = 1024;  // units per artificial row
                if (b.n1 == 1 && b.n2 == 1 && b.n0 >= 2 * W) {
                    Box main = b, tail = b;
                    main.n0 = W, main.n1 = b.n0 / W, main.is1 = main.os1 = W;
                    tail.n0 = b.n0 % W;
                    tail.in_off += main.n1 * W, tail.out_off += main.n1 * W;
                    scaled.push_back(main), owner.push_back((int)n);
                    if (tail.n0 > 0) scaled.push_back(tail), owner.push_back((int)n);
                } else {
                    scaled.push_back(b), owner.push_back((int)n);
                }
            }
            auto mk = [&](const Box& b, long long begin, int n) {
                return make_desc(b, tx, ty * kRowsPerThread, begin, n);
            };
            tx_slot_[slot] = tx;
            all_[slot].offset = (long long)host.size();
            long long item = 0;
            for (size_t i = 0; i < scaled.size(); ++i) {
                BlockDesc d = mk(scaled[i], item, owner[i]);
                long long cnt = (long long)d.tiles0 * d.tiles1 * d.n2;
                if (cnt >= (1ll << 31)) return DTFFTB_ERROR_INTERNAL;
                item += cnt;
                host.push_back(d);
                all_[slot].nblocks++;
            }
            all_[slot].total_items = item;
            for (size_t n = 0; n < boxes_.size(); ++n) {
                long long it = 0;
                single_[slot][n].offset = (long long)host.size();
                for (size_t i = 0; i < scaled.size(); ++i) {
                    if (owner[i] != (int)n) continue;
                    BlockDesc d = mk(scaled[i], it, owner[i]);
                    it += (long long)d.tiles0 * d.tiles1 * d.n2;
                    host.push_back(d);
                    single_[slot][n].nblocks++;
                }
                single_[slot][n].total_items = it;
            }
            if (unit == unit_geo_) tx_ = tx;
        }
        grid_cap_ = grid_mult ? sm_count_ * 8 * grid_mult : 0x7fffffff;
    }
    // interleave the peers of an all-peer table whose blocks live in different GPUs
    bool remote = false;
    for (void* p : peer_out_) remote |= p != nullptr;
    if (remote && getenv("DTFFTB_NO_SHUFFLE") == nullptr) {
        for (int slot = 0; slot < 3; ++slot) {
            const DeviceTable& t = all_[slot];
            if (t.nblocks < 2 || t.total_items < 4 || t.total_items >= (1ll << 31)) continue;
            long long s = (long long)(0.6180339887 * (double)t.total_items);
            auto gcd = [](long long a, long long b) {
                while (b) {
                    long long r = a % b;
                    a = b, b = r;
                }
                return a;
            };
            while (s > 1 && gcd(s, t.total_items) != 1) --s;
            host[(size_t)t.offset].shuffle = s;
        }
    }
    if (abort_flag_) {  // the first block of every table a launch can start from
        for (int slot = 0; slot < 3; ++slot) {
            if (all_[slot].nblocks > 0) host[(size_t)all_[slot].offset].abort = abort_flag_;
            for (const DeviceTable& t : single_[slot])
                if (t.nblocks > 0) host[(size_t)t.offset].abort = abort_flag_;
        }
    }
    if (dry_) {
        host_tables_ = host;
        return DTFFT_SUCCESS;
    }
    if (!host.empty()) {
        cudaError_t ce = cudaMalloc(&d_blocks_, host.size() * sizeof(BlockDesc));
        if (ce != cudaSuccess) return cuda_error(ce);
        ce = cudaMemcpy(d_blocks_, host.data(), host.size() * sizeof(BlockDesc), cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) return cuda_error(ce);
    }
    return DTFFT_SUCCESS;
}

const BlockDesc* Kernel::host_table(int unit, int neighbor, DeviceTable* t, int launch[3]) const {
    if (!dry_ || !created_ || noop_ || (family_ != FAM_T && family_ != FAM_R)) return nullptr;
    int slot = 0;
    if (family_ == FAM_R) {
        if (unit != 4 && unit != 8 && unit != 16) return nullptr;
        if (unit > unit_geo_) return nullptr;
        slot = unit == 4 ? 0 : unit == 8 ? 1 : 2;
        launch[0] = tx_slot_[slot], launch[1] = kRowsThreads / tx_slot_[slot], launch[2] = kRowsPerThread;
    } else {
        launch[0] = tile_.ka, launch[1] = tile_.kb, launch[2] = tile_.rows;
    }
    const DeviceTable* dt = nullptr;
    if (neighbor == 0) {
        dt = &all_[slot];
    } else {
        if (neighbor < 1 || neighbor > (int)single_[slot].size()) return nullptr;
        dt = &single_[slot][(size_t)(neighbor - 1)];
    }
    *t = *dt;
    if (dt->nblocks == 0) return host_tables_.data();  // empty table: nothing to launch
    return host_tables_.data() + dt->offset;
}

int Kernel::pick_unit(const void* in, const void* out) const {
    int u = unit_geo_;
    auto al = [&](const void* p) {
        while (u > 4 && (reinterpret_cast<uintptr_t>(p) % u) != 0) u >>= 1;
    };
    al(in);
    al(out);
    for (void* p : peer_out_)
        if (p) al(p);
    return u;
}

int Kernel::launch(const DeviceTable& t, int unit, const void* in, void* out, cudaStream_t stream) {
    if (t.nblocks == 0 || t.total_items == 0) return DTFFT_SUCCESS;
    cudaError_t ce;
    const int cap = grid_limit_ > 0 ? std::min(grid_cap_, grid_limit_) : grid_cap_;
    if (family_ == FAM_T) {
        // family T moves whole elements with es_-wide accesses: every base it touches must be aligned
        // to the element size (family R narrows its unit instead, pick_unit)
        const uintptr_t mask = (uintptr_t)es_ - 1;
        if ((reinterpret_cast<uintptr_t>(in) & mask) || (reinterpret_cast<uintptr_t>(out) & mask)) return DTFFT_ERROR_INVALID_USAGE;
        bool noshift = (reinterpret_cast<uintptr_t>(out) & 127) != 0;
        for (void* p : peer_out_) {
            if (reinterpret_cast<uintptr_t>(p) & mask) return DTFFT_ERROR_INVALID_USAGE;
            noshift |= (reinterpret_cast<uintptr_t>(p) & 127) != 0;
        }
        ce = launch_transpose((int)es_, tile_, in, out, d_blocks_ + t.offset, t.nblocks, t.total_items, cap, stream, noshift);
    } else {
        const int slot = unit == 4 ? 0 : unit == 8 ? 1 : 2;
        ce = launch_rows(unit, tx_slot_[slot], in, out, d_blocks_ + t.offset, t.nblocks, t.total_items, cap, stream);
    }
    return ce == cudaSuccess ? DTFFT_SUCCESS : cuda_error(ce);
}

int Kernel::execute(const void* in, void* out, cudaStream_t stream, int neighbor, bool sync) {
    if (!created_) return DTFFTB_ERROR_INTERNAL;
    if (dry_) return DTFFT_ERROR_GPU_NOT_SET;
    if (noop_) return DTFFT_SUCCESS;  // abstract_kernel.F90:310
    if (!in || !out) return DTFFT_ERROR_INVALID_USAGE;
    int rc = DTFFT_SUCCESS;
    cudaError_t ce = cudaSuccess;
    if (type_ == K_COPY) {  // kernel_device.F90:117-127
        long long n = (long long)dims_[0] * dims_[1] * dims_[2];
        ce = cudaMemcpyAsync(out, in, (size_t)(n * es_), cudaMemcpyDeviceToDevice, stream);
        if (ce != cudaSuccess) return cuda_error(ce);
    } else if (type_ == K_COPY_PIPELINED) {  // kernel_device.F90:129-139
        if (neighbor < 1 || neighbor > P_) return DTFFTB_ERROR_INTERNAL;
        const int32_t* l = &nd_[5 * (size_t)(neighbor - 1)];
        long long n = (long long)l[0] * l[1] * (ndims_ == 3 ? l[2] : 1);
        if (n > 0) {
            ce = cudaMemcpyAsync(static_cast<char*>(out) + es_ * l[4], static_cast<const char*>(in) + es_ * l[3],
                                 (size_t)(n * es_), cudaMemcpyDeviceToDevice, stream);
            if (ce != cudaSuccess) return cuda_error(ce);
        }
    } else {
        int unit = (int)es_;
        int slot = 0;
        if (family_ != FAM_T) {  // family T: alignment is checked in launch()
            unit = pick_unit(in, out);
            slot = unit == 4 ? 0 : unit == 8 ? 1 : 2;
        }
        if (is_per_neighbor_kind(type_)) {
            if (neighbor < 1 || neighbor > P_) return DTFFTB_ERROR_INTERNAL;  // "Neighbor index out of bounds"
            rc = launch(single_[slot][(size_t)neighbor - 1], unit, in, out, stream);
        } else {
            rc = launch(all_[slot], unit, in, out, stream);
        }
        if (rc) return rc;
    }
    if (sync) {
        ce = cudaStreamSynchronize(stream);
        if (ce != cudaSuccess) return cuda_error(ce);
    }
    return DTFFT_SUCCESS;
}

int Kernel::execute_all(const void* in, void* out, cudaStream_t stream) {
    if (!created_) return DTFFTB_ERROR_INTERNAL;
    if (dry_) return DTFFT_ERROR_GPU_NOT_SET;
    if (noop_) return DTFFT_SUCCESS;
    if (family_ == FAM_COPY) {
        if (type_ == K_COPY) return execute(in, out, stream, 0, false);
        for (int n = 1; n <= P_; ++n) {
            int rc = execute(in, out, stream, n, false);
            if (rc) return rc;
        }
        return DTFFT_SUCCESS;
    }
    if (!in || !out) return DTFFT_ERROR_INVALID_USAGE;
    int unit = (int)es_, slot = 0;
    if (family_ == FAM_R) {
        unit = pick_unit(in, out);
        slot = unit == 4 ? 0 : unit == 8 ? 1 : 2;
    }
    return launch(all_[slot], unit, in, out, stream);
}

int Kernel::set_peer_out(void* const* out_bases, const int64_t* out_displs_override) {
    if (!created_) return DTFFTB_ERROR_INTERNAL;
    if (noop_ || family_ == FAM_COPY) return DTFFT_SUCCESS;
    if (P_ == 0) return DTFFTB_ERROR_INTERNAL;
    peer_out_.clear();
    peer_out_displ_.clear();
    const int ptype = per_neighbor_type(type_);
    if (!custom_)
        for (int n = 0; n < P_; ++n) boxes_[n] = make_box(ptype, ndims_, dims_, &nd_[5 * (size_t)n]);
    if (out_bases) {
        peer_out_.assign(out_bases, out_bases + P_);
        if (out_displs_override) {
            peer_out_displ_.assign(out_displs_override, out_displs_override + P_);
            for (int n = 0; n < P_; ++n) boxes_[n].out_off = out_displs_override[n];
        }
    }
    return rebuild_tables();
}

int Kernel::set_tile(int ka, int kb, int rows) {
    if (!created_) return DTFFTB_ERROR_INTERNAL;
    if (noop_ || family_ != FAM_T) return DTFFT_SUCCESS;
    if (!transpose_cfg_supported((int)es_, TileCfg{ka, kb, rows})) return DTFFT_ERROR_INVALID_USAGE;
    tile_ = TileCfg{ka, kb, rows};
    return rebuild_tables();
}

int Kernel::autotune(const void* in, void* out, cudaStream_t stream, int n_warmup, int n_iters, float* best_ms) {
    if (best_ms) *best_ms = 0.f;
    if (!created_) return DTFFTB_ERROR_INTERNAL;
    if (noop_ || family_ != FAM_T) return DTFFT_SUCCESS;
    static const TileCfg cands[] = {{1, 1, 4}, {1, 1, 8}, {1, 1, 16}, {2, 1, 8}, {1, 2, 8},
                                    {2, 2, 8}, {2, 2, 16}, {1, 4, 16}, {4, 1, 16}};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    TileCfg best_cfg = tile_;
    for (const TileCfg& c : cands) {
        if (set_tile(c.ka, c.kb, c.rows) != DTFFT_SUCCESS) continue;
        int rc = 0;
        for (int i = 0; i < n_warmup && !rc; ++i) rc = execute_all(in, out, stream);
        cudaEventRecord(e0, stream);
        for (int i = 0; i < n_iters && !rc; ++i) rc = execute_all(in, out, stream);
        cudaEventRecord(e1, stream);
        if (cudaEventSynchronize(e1) != cudaSuccess || rc) {
            cudaGetLastError();
            continue;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= std::max(1, n_iters);
        if (getenv("DTFFTB_LOG"))  // kernel_device.F90:385-389 prints time and bandwidth of every candidate
            fprintf(stderr, "[dtfftb] autotune es=%d tile=%dx%d rows=%d: %.4f ms, %.1f GB/s\n", (int)es_, 32 * c.ka,
                    32 * c.kb, c.rows, ms, ms > 0 ? 2.0 * (double)bytes_moved() / (ms * 1e-3) / 1e9 : 0.0);
        if (autotune_log_) autotune_log_->push_back(AutotuneEntry{c, ms, ms > 0 ? 2.0 * (double)bytes_moved() / (ms * 1e-3) / 1e9 : 0.0});
        if (ms < best) best = ms, best_cfg = c;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (best_ms) *best_ms = best;
    return set_tile(best_cfg.ka, best_cfg.kb, best_cfg.rows);
}

long long Kernel::csize(int neighbor) const {
    if (neighbor < 1 || neighbor > P_) return 0;
    const int32_t* l = &nd_[5 * (size_t)(neighbor - 1)];
    return (long long)l[0] * l[1] * l[2] * (es_ / 4);
}

long long Kernel::bytes_moved() const {
    if (noop_) return 0;
    if (family_ == FAM_COPY && type_ == K_COPY) return (long long)dims_[0] * dims_[1] * dims_[2] * es_;
    long long e = 0;
    for (auto& b : boxes_)
        if (!b.empty()) e += b.volume();
    return e * es_;
}

void Kernel::get_info(int* family, int* unit, int* tile_a, int* tile_b, int* threads, int64_t* n_items) const {
    if (family) *family = noop_ ? FAM_NONE : family_;
    if (unit) *unit = family_ == FAM_R ? unit_geo_ : (int)es_;
    if (tile_a) *tile_a = family_ == FAM_T ? 32 * tile_.ka : (family_ == FAM_R ? tx_ : 0);
    if (tile_b) *tile_b = family_ == FAM_T ? 32 * tile_.kb : (family_ == FAM_R ? (kRowsThreads / tx_) * kRowsPerThread : 0);
    if (threads) *threads = family_ == FAM_T ? 32 * tile_.rows : (family_ == FAM_R ? kRowsThreads : 0);
    if (n_items) {
        const int slot = family_ == FAM_R ? (unit_geo_ == 4 ? 0 : unit_geo_ == 8 ? 1 : 2) : 0;
        *n_items = noop_ ? 0 : all_[slot].total_items;
    }
}

}  // namespace dtfftb
