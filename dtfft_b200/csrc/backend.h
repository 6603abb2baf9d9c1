// Exchange backends of the reshape path.
//
//   NcclBackend   -- replaces backend_nccl (src/dtfft_backend_nccl.F90:65-134) behind the
//                    contract of abstract_backend (src/dtfft_abstract_backend.F90:143-343):
//                    grouped ncclSend/ncclRecv all-to-all(v) in 4-byte float units on the plan
//                    stream; the pipelined flavour copies the self block on a second stream
//                    and unpacks peer by peer.
//   The NVLink direct-store exchange has no separate backend object: it is one fused kernel
//   bracketed by two device barriers and lives in ReshapeHandle (handle.cu).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>
#include <vector>

#include "../../include/dtfft_b200_api.h"
#include "kernel_object.h"

namespace dtfftb {

// dtfft_backend_t values on this path (include/dtfft_config.h.in:153-169) + the new one.
enum BackendType : int {
    BACKEND_NONE = -111,
    BACKEND_NCCL = 24,
    BACKEND_NCCL_PIPELINED = 27,
    BACKEND_NVLINK_FUSED = 38,  // new: fused pack + remote store over NVLink peer memory
};

inline bool backend_is_pipelined(int b) { return b == BACKEND_NCCL_PIPELINED; }

inline int nccl_error(ncclResult_t r) { return r == ncclSuccess ? 0 : DTFFTB_ERROR_NCCL_BASE - (int)r; }

class NcclBackend {
public:
    ~NcclBackend() { destroy(); }
    // counts / displs in ELEMENTS of `base_storage` bytes, one entry per member of the 1-D
    // communicator; `mapping[i]` = rank of member i in `nccl` (comm_mappings,
    // src/dtfft_abstract_backend.F90:432-441).
    int create(int backend, ncclComm_t nccl, int comm_rank, const std::vector<int>& mapping,
               const std::vector<int64_t>& send_displs, const std::vector<int64_t>& send_counts,
               const std::vector<int64_t>& recv_displs, const std::vector<int64_t>& recv_counts, int64_t base_storage);
    void set_unpack_kernel(Kernel* k) { unpack_ = k; }
    bool has_unpack_kernel() const { return unpack_ != nullptr; }
    // abstract_backend%execute: `in` packed send buffer, `out` final destination (pipelined)
    // or receive buffer (plain), `aux` receive workspace of the pipelined flavour.
    int execute(void* in, void* out, cudaStream_t stream, void* aux);
    int64_t aux_bytes() const { return aux_bytes_; }
    bool is_pipelined() const { return pipelined_; }
    void destroy();

private:
    int backend_ = BACKEND_NCCL;
    bool pipelined_ = false;
    ncclComm_t nccl_ = nullptr;
    int P_ = 0, me_ = 0;
    std::vector<int> mapping_;
    std::vector<int64_t> sdispl_, sfloats_, rdispl_, rfloats_;  // float units, displs 0-based
    int64_t self_sdispl_ = 0, self_rdispl_ = 0, self_bytes_ = 0;
    int64_t aux_bytes_ = 0;
    Kernel* unpack_ = nullptr;
    cudaStream_t copy_stream_ = nullptr;
    cudaEvent_t exec_event_ = nullptr, copy_event_ = nullptr;
};

}  // namespace dtfftb
