// C ABI shell over dtfftb::Kernel (declared in include/dtfft_b200.h).
#include <cuda_runtime.h>

#include <new>

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

}  // extern "C"
