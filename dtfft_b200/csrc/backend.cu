// NCCL exchange backend (see backend.h for the reference citations).
#include "backend.h"

#include <algorithm>

#include "errors.h"

namespace dtfftb {

int NcclBackend::create(int backend, ncclComm_t nccl, int comm_rank, const std::vector<int>& mapping,
                        const std::vector<int64_t>& send_displs, const std::vector<int64_t>& send_counts,
                        const std::vector<int64_t>& recv_displs, const std::vector<int64_t>& recv_counts,
                        int64_t base_storage) {
    destroy();
    backend_ = backend;
    pipelined_ = backend_is_pipelined(backend);
    nccl_ = nccl;
    me_ = comm_rank;
    mapping_ = mapping;
    P_ = (int)mapping.size();
    // abstract_backend.F90:160-183: everything is expressed in 4-byte floats
    const int64_t scaler = base_storage / 4;
    sdispl_.resize(P_), sfloats_.resize(P_), rdispl_.resize(P_), rfloats_.resize(P_);
    int64_t ssum = 0, rsum = 0;
    for (int i = 0; i < P_; ++i) {
        sdispl_[i] = send_displs[i] * scaler, sfloats_[i] = send_counts[i] * scaler;
        rdispl_[i] = recv_displs[i] * scaler, rfloats_[i] = recv_counts[i] * scaler;
        ssum += sfloats_[i], rsum += rfloats_[i];
    }
    aux_bytes_ = pipelined_ ? std::max(ssum, rsum) * 4 : 0;  // :196-201
    self_bytes_ = 0;
    if (pipelined_) {  // is_selfcopy, :186-215
        self_sdispl_ = sdispl_[me_], self_rdispl_ = rdispl_[me_];
        self_bytes_ = sfloats_[me_] * 4;
        sfloats_[me_] = 0, rfloats_[me_] = 0;
        cudaError_t ce = cudaEventCreateWithFlags(&exec_event_, cudaEventDisableTiming);
        if (ce != cudaSuccess) return cuda_error(ce);
        ce = cudaEventCreateWithFlags(&copy_event_, cudaEventDisableTiming);
        if (ce != cudaSuccess) return cuda_error(ce);
        ce = cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking);
        if (ce != cudaSuccess) return cuda_error(ce);
    }
    return DTFFT_SUCCESS;
}

int NcclBackend::execute(void* in, void* out, cudaStream_t stream, void* aux) {
    float* pin = static_cast<float*>(in);
    float* pout = pipelined_ ? static_cast<float*>(aux) : static_cast<float*>(out);
    if (pipelined_ && !aux) return DTFFTB_ERROR_INTERNAL;
    cudaError_t ce;
    int rc;
    if (pipelined_ && self_bytes_ > 0) {  // abstract_backend.F90:269-277, 303-331
        ce = cudaEventRecord(exec_event_, stream);
        if (ce != cudaSuccess) return cuda_error(ce);
        ce = cudaStreamWaitEvent(copy_stream_, exec_event_, 0);
        if (ce != cudaSuccess) return cuda_error(ce);
        ce = cudaMemcpyAsync(pout + self_rdispl_, pin + self_sdispl_, (size_t)self_bytes_, cudaMemcpyDeviceToDevice,
                             copy_stream_);
        if (ce != cudaSuccess) return cuda_error(ce);
        if (unpack_) {  // null = unpack-free reshape: the block is already in place
            rc = unpack_->execute(aux, out, copy_stream_, me_ + 1, false);
            if (rc) return rc;
        }
    }
    // backend_nccl.F90:108-119
    ncclResult_t nr = ncclGroupStart();
    if (nr != ncclSuccess) return nccl_error(nr);
    for (int i = 0; i < P_; ++i) {
        if (i == me_ && pipelined_) continue;
        const int rnk = mapping_[i];
        if (sfloats_[i] > 0) {
            nr = ncclSend(pin + sdispl_[i], (size_t)sfloats_[i], ncclFloat, rnk, nccl_, stream);
            if (nr != ncclSuccess) break;
        }
        if (rfloats_[i] > 0) {
            nr = ncclRecv(pout + rdispl_[i], (size_t)rfloats_[i], ncclFloat, rnk, nccl_, stream);
            if (nr != ncclSuccess) break;
        }
    }
    // the group is closed on the error path too: an open group would swallow every later NCCL call
    const ncclResult_t ne = ncclGroupEnd();
    if (nr != ncclSuccess) return nccl_error(nr);
    if (ne != ncclSuccess) return nccl_error(ne);
    if (pipelined_) {  // :126-132
        for (int i = 0; i < P_; ++i) {
            if (rfloats_[i] > 0 && unpack_) {
                rc = unpack_->execute(aux, out, stream, i + 1, false);
                if (rc) return rc;
            }
        }
        if (self_bytes_ > 0) {  // wait(), abstract_backend.F90:333-343
            ce = cudaEventRecord(copy_event_, copy_stream_);
            if (ce != cudaSuccess) return cuda_error(ce);
            ce = cudaStreamWaitEvent(stream, copy_event_, 0);
            if (ce != cudaSuccess) return cuda_error(ce);
        }
    }
    return DTFFT_SUCCESS;
}

void NcclBackend::destroy() {
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    if (exec_event_) cudaEventDestroy(exec_event_);
    if (copy_event_) cudaEventDestroy(copy_event_);
    copy_stream_ = nullptr;
    exec_event_ = copy_event_ = nullptr;
    unpack_ = nullptr;
    nccl_ = nullptr;
}

}  // namespace dtfftb
