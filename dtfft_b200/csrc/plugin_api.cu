// C ABI shells over the exchange backend and the FFT executor (declared in
// include/dtfft_b200.h): what a Fortran `backend_nccl` / `cufft_executor` replacement binds with
// iso_c_binding when dtFFT keeps its own plan and reshape handles (level 1).
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstring>
#include <new>
#include <vector>

#include "../../include/dtfft_b200.h"
#include "backend.h"
#include "comm.h"
#include "errors.h"
#include "c_handles.h"
#include "kernel_object.h"
#include "plan.h"

struct dtfftb_backend_s {
    dtfftb::NcclBackend b;
};
struct dtfftb_executor_s {
    dtfftb::FftExecutor f;
};

extern "C" {

int dtfftb_nccl_comm_create(const dtfftb_comm_t* comm, void** nccl_comm) {
    // backend_helper%create, src/dtfft_abstract_backend.F90:432-457: rank 0 makes the id, everybody
    // learns it (MPI_Bcast there, the allgather callback here), ncclCommInitRank over the group
    if (!nccl_comm) return DTFFT_ERROR_INVALID_USAGE;
    *nccl_comm = nullptr;
    dtfftb::Comm c(comm);
    ncclUniqueId id;
    std::memset(&id, 0, sizeof(id));
    if (c.rank() == 0) {
        ncclResult_t nr = ncclGetUniqueId(&id);
        if (nr != ncclSuccess) return dtfftb::nccl_error(nr);
    }
    std::vector<ncclUniqueId> all;
    if (c.allgather_v(id, all)) return DTFFTB_ERROR_COMM;
    ncclComm_t out = nullptr;
    ncclResult_t nr = ncclCommInitRank(&out, c.size(), all[0], c.rank());
    if (nr != ncclSuccess) return dtfftb::nccl_error(nr);
    *nccl_comm = out;
    return DTFFT_SUCCESS;
}

int dtfftb_nccl_comm_destroy(void** nccl_comm) {
    if (!nccl_comm || !*nccl_comm) return DTFFT_ERROR_INVALID_USAGE;
    ncclResult_t nr = ncclCommDestroy(static_cast<ncclComm_t>(*nccl_comm));
    *nccl_comm = nullptr;
    return dtfftb::nccl_error(nr);
}

int dtfftb_backend_create(dtfftb_backend_t* backend, int backend_type, void* nccl_comm, int comm_rank, int comm_size,
                          const int32_t* comm_mapping, const int64_t* send_displs, const int64_t* send_counts,
                          const int64_t* recv_displs, const int64_t* recv_counts, int64_t base_storage) {
    if (!backend) return DTFFT_ERROR_INVALID_USAGE;
    *backend = nullptr;
    if (backend_type != dtfftb::BACKEND_NCCL && backend_type != dtfftb::BACKEND_NCCL_PIPELINED) return DTFFT_ERROR_INVALID_BACKEND;
    if (!nccl_comm || comm_size < 1 || comm_rank < 0 || comm_rank >= comm_size || !send_displs || !send_counts ||
        !recv_displs || !recv_counts)
        return DTFFT_ERROR_INVALID_USAGE;
    if (base_storage != 4 && base_storage != 8 && base_storage != 16) return DTFFT_ERROR_INVALID_USAGE;
    dtfftb_backend_s* h = new (std::nothrow) dtfftb_backend_s;
    if (!h) return DTFFT_ERROR_ALLOC_FAILED;
    std::vector<int> mapping((size_t)comm_size);
    for (int i = 0; i < comm_size; ++i) mapping[(size_t)i] = comm_mapping ? comm_mapping[i] : i;
    auto vec = [&](const int64_t* p) { return std::vector<int64_t>(p, p + comm_size); };
    int rc = h->b.create(backend_type, static_cast<ncclComm_t>(nccl_comm), comm_rank, mapping, vec(send_displs),
                         vec(send_counts), vec(recv_displs), vec(recv_counts), base_storage);
    if (rc != DTFFT_SUCCESS) {
        delete h;
        return rc;
    }
    *backend = h;
    return DTFFT_SUCCESS;
}

int dtfftb_backend_set_unpack_kernel(dtfftb_backend_t backend, dtfftb_kernel_t unpack_kernel) {
    if (!backend) return DTFFT_ERROR_INVALID_USAGE;
    backend->b.set_unpack_kernel(unpack_kernel ? &unpack_kernel->k : nullptr);
    return DTFFT_SUCCESS;
}

int dtfftb_backend_get_aux_bytes(dtfftb_backend_t backend, int64_t* aux_bytes) {
    if (!backend || !aux_bytes) return DTFFT_ERROR_INVALID_USAGE;
    *aux_bytes = backend->b.aux_bytes();
    return DTFFT_SUCCESS;
}

int dtfftb_backend_execute(dtfftb_backend_t backend, void* in, void* out, void* stream, void* aux) {
    if (!backend || !in || !out) return DTFFT_ERROR_INVALID_USAGE;
    if (backend->b.is_pipelined() && !aux) return DTFFT_ERROR_INVALID_AUX;
    // the pipelined flavour unpacks block by block: a caller with nothing to unpack (unpack-free
    // reshape) hands it a KERNEL_DUMMY kernel like the reference does (reshape_handle_generic.F90:623)
    if (backend->b.is_pipelined() && !backend->b.has_unpack_kernel()) return DTFFT_ERROR_INVALID_USAGE;
    return backend->b.execute(in, out, static_cast<cudaStream_t>(stream), aux);
}

int dtfftb_backend_destroy(dtfftb_backend_t* backend) {
    if (!backend || !*backend) return DTFFT_ERROR_INVALID_USAGE;
    delete *backend;
    *backend = nullptr;
    return DTFFT_SUCCESS;
}

int dtfftb_executor_create(dtfftb_executor_t* executor, int fft_rank, int fft_type, int precision, int32_t idist,
                           int32_t odist, int32_t how_many, const int32_t* fft_sizes, const int32_t* inembed,
                           const int32_t* onembed, void* stream) {
    if (!executor) return DTFFT_ERROR_INVALID_USAGE;
    *executor = nullptr;
    if (fft_rank != 1 && fft_rank != 2) return DTFFT_ERROR_INVALID_USAGE;
    if (fft_type == 2) return DTFFT_ERROR_R2R_FFT_NOT_SUPPORTED;  // cuFFT has no r2r (dtfft_executor_cufft_m.F90:94-98)
    if (fft_type != 0 && fft_type != 1) return DTFFT_ERROR_INVALID_USAGE;
    if (precision != DTFFT_SINGLE && precision != DTFFT_DOUBLE) return DTFFT_ERROR_INVALID_PRECISION;
    if (!fft_sizes || !inembed || !onembed) return DTFFT_ERROR_INVALID_USAGE;
    dtfftb_executor_s* h = new (std::nothrow) dtfftb_executor_s;
    if (!h) return DTFFT_ERROR_ALLOC_FAILED;
    int n[2] = {1, 1}, ie[2] = {1, 1}, oe[2] = {1, 1};
    for (int i = 0; i < fft_rank; ++i) n[i] = fft_sizes[i], ie[i] = inembed[i], oe[i] = onembed[i];
    int rc = h->f.create_raw(fft_rank, fft_type == 1, precision, idist, odist, how_many, n, ie, oe,
                             static_cast<cudaStream_t>(stream));
    if (rc != DTFFT_SUCCESS) {
        delete h;
        return rc;
    }
    *executor = h;
    return DTFFT_SUCCESS;
}

int dtfftb_executor_execute(dtfftb_executor_t executor, void* a, void* b, int sign) {
    if (!executor || !a || !b || (sign != -1 && sign != 1)) return DTFFT_ERROR_INVALID_USAGE;
    return executor->f.execute(a, b, sign);
}

int dtfftb_executor_destroy(dtfftb_executor_t* executor) {
    if (!executor || !*executor) return DTFFT_ERROR_INVALID_USAGE;
    delete *executor;
    *executor = nullptr;
    return DTFFT_SUCCESS;
}

}  // extern "C"
