// C ABI shell over dtfftb::Kernel (declared in include/dtfft_b200.h).
#include <cuda_runtime.h>

#include <cstdint>
#include <new>
#include <vector>

#include "../../include/dtfft_b200.h"
#include "errors.h"
#include "c_handles.h"
#include "kernel_object.h"


extern "C" {

const char* dtfftb_version(void) { return "dtfft_b200 0.1.0 (dtFFT 3.2.0 reshape path, sm_100a)"; }

int dtfftb_device_available(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n > 0 ? 1 : 0;
}

int dtfftb_kernel_create(dtfftb_kernel_t* kernel, int ndims, const int32_t* dims, int kernel_type,
                         int64_t base_storage, const int32_t* neighbor_data, int n_neighbors, int effort,
                         int force_effort) {
    if (!kernel) return DTFFT_ERROR_INVALID_USAGE;
    *kernel = nullptr;
    dtfftb_kernel_s* h = new (std::nothrow) dtfftb_kernel_s;
    if (!h) return DTFFT_ERROR_ALLOC_FAILED;
    int rc = h->k.create(ndims, dims, kernel_type, base_storage, neighbor_data, n_neighbors, effort, force_effort != 0);
    if (rc != DTFFT_SUCCESS) {
        delete h;
        return rc;
    }
    *kernel = h;
    return DTFFT_SUCCESS;
}

int dtfftb_kernel_create_dry(dtfftb_kernel_t* kernel, int ndims, const int32_t* dims, int kernel_type,
                             int64_t base_storage, const int32_t* neighbor_data, int n_neighbors) {
    if (!kernel) return DTFFT_ERROR_INVALID_USAGE;
    *kernel = nullptr;
    dtfftb_kernel_s* h = new (std::nothrow) dtfftb_kernel_s;
    if (!h) return DTFFT_ERROR_ALLOC_FAILED;
    h->k.set_dry(true);
    int rc = h->k.create(ndims, dims, kernel_type, base_storage, neighbor_data, n_neighbors, 0, false);
    if (rc != DTFFT_SUCCESS) {
        delete h;
        return rc;
    }
    *kernel = h;
    return DTFFT_SUCCESS;
}

int dtfftb_kernel_create_boxes_dry(dtfftb_kernel_t* kernel, int family, int64_t base_storage, int n_boxes,
                                   const int64_t* boxes, int remote_peers) {
    if (!kernel || !boxes || n_boxes <= 0) return DTFFT_ERROR_INVALID_USAGE;
    *kernel = nullptr;
    dtfftb_kernel_s* h = new (std::nothrow) dtfftb_kernel_s;
    if (!h) return DTFFT_ERROR_ALLOC_FAILED;
    h->k.set_dry(true);
    std::vector<dtfftb::Box> bx((size_t)n_boxes);
    for (int i = 0; i < n_boxes; ++i) {
        const int64_t* o = boxes + 10 * i;
        dtfftb::Box& b = bx[(size_t)i];
        b.n0 = o[0], b.n1 = o[1], b.n2 = o[2], b.in_off = o[3], b.out_off = o[4];
        b.is1 = o[5], b.is2 = o[6], b.os0 = o[7], b.os1 = o[8], b.os2 = o[9];
    }
    int rc = h->k.create_boxes((dtfftb::Family)family, base_storage, bx);
    if (rc == DTFFT_SUCCESS && remote_peers && !h->k.is_noop()) {
        // stand-ins for peer-mapped destination bases: box i is written to "peer i + 1"
        std::vector<void*> bases((size_t)n_boxes);
        for (int i = 0; i < n_boxes; ++i) bases[(size_t)i] = reinterpret_cast<void*>((uintptr_t)(i + 1) << 40);
        rc = h->k.set_peer_out(bases.data(), nullptr);
    }
    if (rc != DTFFT_SUCCESS) {
        delete h;
        return rc;
    }
    *kernel = h;
    return DTFFT_SUCCESS;
}

int dtfftb_kernel_create_boxes(dtfftb_kernel_t* kernel, int family, int64_t base_storage, int n_boxes,
                               const int64_t* boxes, void* const* out_bases) {
    if (!kernel || !boxes || n_boxes <= 0) return DTFFT_ERROR_INVALID_USAGE;
    *kernel = nullptr;
    dtfftb_kernel_s* h = new (std::nothrow) dtfftb_kernel_s;
    if (!h) return DTFFT_ERROR_ALLOC_FAILED;
    std::vector<dtfftb::Box> bx((size_t)n_boxes);
    for (int i = 0; i < n_boxes; ++i) {
        const int64_t* o = boxes + 10 * i;
        dtfftb::Box& b = bx[(size_t)i];
        b.n0 = o[0], b.n1 = o[1], b.n2 = o[2], b.in_off = o[3], b.out_off = o[4];
        b.is1 = o[5], b.is2 = o[6], b.os0 = o[7], b.os1 = o[8], b.os2 = o[9];
    }
    int rc = h->k.create_boxes((dtfftb::Family)family, base_storage, bx);
    if (rc == DTFFT_SUCCESS && out_bases && !h->k.is_noop()) rc = h->k.set_peer_out(out_bases, nullptr);
    if (rc != DTFFT_SUCCESS) {
        delete h;
        return rc;
    }
    *kernel = h;
    return DTFFT_SUCCESS;
}

int dtfftb_kernel_dump_table(dtfftb_kernel_t kernel, int unit, int neighbor, int32_t cap, int64_t* rows,
                             int32_t* n_blocks, int64_t* total_items, int32_t* launch) {
    if (!kernel || !n_blocks || !total_items || !launch) return DTFFT_ERROR_INVALID_USAGE;
    if (!kernel->k.dry()) return DTFFT_ERROR_INVALID_USAGE;
    *n_blocks = 0, *total_items = 0;
    dtfftb::DeviceTable t;
    int l3[3] = {0, 0, 0};
    const dtfftb::BlockDesc* b = kernel->k.host_table(unit, neighbor, &t, l3);
    if (!b) return DTFFT_SUCCESS;  // no such table (no-op kernel, copy kernel, unit too wide)
    *n_blocks = t.nblocks, *total_items = t.total_items;
    for (int i = 0; i < 3; ++i) launch[i] = l3[i];
    if (!rows || cap < t.nblocks) return DTFFT_SUCCESS;
    for (int i = 0; i < t.nblocks; ++i) {
        const dtfftb::BlockDesc& d = b[i];
        int64_t* o = rows + 20 * i;
        o[0] = d.in_off, o[1] = d.out_off, o[2] = d.is1, o[3] = d.is2, o[4] = d.os0, o[5] = d.os1, o[6] = d.os2;
        o[7] = d.item_begin, o[8] = d.shuffle, o[9] = d.n0, o[10] = d.n1, o[11] = d.n2, o[12] = d.tiles0, o[13] = d.tiles1;
        o[14] = d.div0.mul, o[15] = d.div0.shr, o[16] = d.div1.mul, o[17] = d.div1.shr;
        o[18] = d.out_base ? (int64_t)((uintptr_t)d.out_base >> 40) - 1 : -1;
        o[19] = d.bshift;
    }
    return DTFFT_SUCCESS;
}

int dtfftb_kernel_execute(dtfftb_kernel_t kernel, const void* in, void* out, void* stream, int neighbor, int sync) {
    if (!kernel) return DTFFT_ERROR_INVALID_USAGE;
    return kernel->k.execute(in, out, static_cast<cudaStream_t>(stream), neighbor, sync != 0);
}

int dtfftb_kernel_execute_all(dtfftb_kernel_t kernel, const void* in, void* out, void* stream) {
    if (!kernel) return DTFFT_ERROR_INVALID_USAGE;
    return kernel->k.execute_all(in, out, static_cast<cudaStream_t>(stream));
}

int dtfftb_kernel_set_peer_out(dtfftb_kernel_t kernel, void* const* out_bases, const int64_t* out_displs_override) {
    if (!kernel) return DTFFT_ERROR_INVALID_USAGE;
    return kernel->k.set_peer_out(out_bases, out_displs_override);
}

int dtfftb_kernel_destroy(dtfftb_kernel_t* kernel) {
    if (!kernel) return DTFFT_ERROR_INVALID_USAGE;
    if (*kernel) delete *kernel;
    *kernel = nullptr;
    return DTFFT_SUCCESS;
}

int dtfftb_kernel_get_info(dtfftb_kernel_t kernel, int* family, int* unit_bytes, int* tile_a, int* tile_b,
                           int* threads, int64_t* n_items) {
    if (!kernel) return DTFFT_ERROR_INVALID_USAGE;
    kernel->k.get_info(family, unit_bytes, tile_a, tile_b, threads, n_items);
    return DTFFT_SUCCESS;
}

int dtfftb_kernel_set_tile(dtfftb_kernel_t kernel, int ka, int kb, int rows) {
    if (!kernel) return DTFFT_ERROR_INVALID_USAGE;
    return kernel->k.set_tile(ka, kb, rows);
}

int dtfftb_kernel_autotune(dtfftb_kernel_t kernel, const void* in, void* out, void* stream, int n_warmup,
                           int n_iters, float* best_ms) {
    if (!kernel) return DTFFT_ERROR_INVALID_USAGE;
    return kernel->k.autotune(in, out, static_cast<cudaStream_t>(stream), n_warmup, n_iters, best_ms);
}

int dtfftb_kernel_autotune_report(dtfftb_kernel_t kernel, const void* in, void* out, void* stream, int n_warmup,
                                  int n_iters, int max_entries, int* n_entries, int32_t* tiles, float* ms, double* gbs) {
    if (!kernel || !n_entries || max_entries < 0) return DTFFT_ERROR_INVALID_USAGE;
    std::vector<dtfftb::Kernel::AutotuneEntry> log;
    kernel->k.set_autotune_log(&log);
    float best = 0.f;
    const int rc = kernel->k.autotune(in, out, static_cast<cudaStream_t>(stream), n_warmup, n_iters, &best);
    kernel->k.set_autotune_log(nullptr);
    *n_entries = (int)log.size();
    for (int i = 0; i < (int)log.size() && i < max_entries; ++i) {
        if (tiles) tiles[3 * i] = 32 * log[(size_t)i].cfg.ka, tiles[3 * i + 1] = 32 * log[(size_t)i].cfg.kb, tiles[3 * i + 2] = 32 * log[(size_t)i].cfg.rows;
        if (ms) ms[i] = log[(size_t)i].ms;
        if (gbs) gbs[i] = log[(size_t)i].gbs;
    }
    return rc;
}

}  // extern "C"
