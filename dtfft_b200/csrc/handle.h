// One transposition / reshape between two decompositions: the B200 replacement of
// reshape_handle_generic (src/dtfft_reshape_handle_generic.F90:174-777).
//
// Three execution shapes:
//   * no exchange (1 rank in the 1-D communicator): a single local permute / copy (:246-253);
//   * NCCL / NCCL_PIPELINED: pack -> all-to-all(v) -> unpack with the reference's buffer
//     choreography (:695-759); transposes use the reference's neighbor_data geometry
//     (geometry.cu), brick reshapes use global-index intersections;
//   * NVLINK_FUSED: ONE kernel reads the local array and stores every element at its final
//     position in the owning peer's `out` through NVLink peer mappings, bracketed by two
//     device barriers -- pack, exchange and unpack of the reference collapse into one pass
//     (1 HBM read + 1 remote/local write per element instead of 3 + 3).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <map>
#include <memory>
#include <vector>

#include "backend.h"
#include "comm.h"
#include "geometry.h"
#include "kernel_object.h"
#include "peer.h"

namespace dtfftb {

struct HandleContext {
    ncclComm_t nccl = nullptr;      // communicator over the whole process grid (may be null when P == 1)
    PeerRegistry* peers = nullptr;  // symmetric-buffer registry (NVLINK_FUSED)
    int effort = 0;
    bool no_shortcuts = false;      // brick reshapes: keep the three-step schedule (no aux needed)
};

class ReshapeHandle {
public:
    ~ReshapeHandle() { destroy(); }
    // `send_by_member[i]` / `recv_by_member[i]`: source / destination layout of member i of the
    // 1-D communicator `members` (world ranks); `me` = my index in it.  `ttype` != 0 for a
    // transposition (dtfft_transpose_t), else `rtype` (dtfft_reshape_t).
    int create(const HandleContext& ctx, int ttype, int rtype, int comm_id, const std::vector<int>& members, int me,
               const std::vector<Pencil>& send_by_member, const std::vector<Pencil>& recv_by_member,
               int64_t base_storage, int backend);
    int execute(void* in, void* out, cudaStream_t stream, void* aux);
    int64_t aux_bytes() const { return aux_bytes_; }
    bool aux_needed() const { return aux_bytes_ > 0; }
    int backend() const { return backend_; }
    bool has_exchange() const { return has_exchange_; }
    // Algorithmic payload of one execute on this rank: elements moved, and those that leave the GPU.
    long long local_elements() const { return send_elems_; }
    long long remote_elements() const { return remote_elems_; }
    int kernel_launches() const { return launches_; }  // of OUR kernels per execute
    const HandleGeometry& geometry() const { return geo_; }
    void destroy();
    // Drop the fused kernels cached per destination address (their peer mappings die with the buffer).
    void forget_buffers() {
        fused_.clear();
        fused_chunks_.clear();
        for (auto& kv : maps_) ctx_.peers->release(&kv.second.opened);
        maps_.clear();
        ++evictions_;
    }
    // Identity check of a destination used before (see peer_bases): true if `out` is unknown to this handle
    // or still the allocation whose peer mappings are cached.
    bool destination_current(const void* out) const {
        auto it = maps_.find(out);
        return it == maps_.end() || it->second.id == buffer_id(out);
    }
    // Bumped whenever cached fused kernels are destroyed: a CUDA graph captured earlier may still
    // point at their device tables and must not be replayed (Plan::execute compares the counters).
    long long evictions() const { return evictions_; }

    // ---- stage overlap on the NVLINK_FUSED path (no reference counterpart) ----------------
    // The source pencil is cut along its SLOWEST axis into `nchunks` ranges; chunk k can be
    // stored to the peers as soon as the producer (an FFT over the same range) has finished,
    // on another stream, while the producer works on chunk k+1:
    //     fused_begin(out, s)                      "every member's `out` is free" barrier
    //     fused_chunk(in, out, k, nchunks, s2) ... persistent kernel limited to `max_ctas` CTAs
    //     fused_end(s)                             "all blocks have landed" barrier
    bool can_chunk() const { return created_ && has_exchange_ && backend_ == BACKEND_NVLINK_FUSED && is_transpose_; }
    long long slow_extent() const { return send_.ndims > 0 ? send_.counts[send_.ndims - 1] : 1; }
    int fused_begin(void* out, cudaStream_t stream);
    int fused_chunk(void* in, void* out, int k, int nchunks, int max_ctas, cudaStream_t stream);
    int fused_end(cudaStream_t stream);

    // A handle without exchange whose kernel is a tiled transpose can run in pieces cut by the blocks of the exchange
    // next to it (local_piece below).
    bool is_local_transpose() const { return created_ && !has_exchange_ && is_transpose_ && send_elems_ > 0; }
    const std::vector<Pencil>& send_by_member() const { return send_by_member_; }
    // Smallest slowest-axis extent of the members' source pencils, 0 if any member's source or
    // destination pencil is empty: what every member can compute alike to agree on a chunk count.
    long long min_member_slow_extent() const {
        long long m = -1;
        for (size_t i = 0; i < send_by_member_.size(); ++i) {
            const Pencil& sp = send_by_member_[i];
            if (sp.size() == 0 || recv_by_member_[i].size() == 0) return 0;
            const long long n = sp.counts[sp.ndims - 1];
            m = m < 0 ? n : std::min(m, n);
        }
        return m < 0 ? 0 : m;
    }

    // ---- NVLINK_FUSED, copy-engine form of the exchange (geometry.h: DmaBlock) ----------------------------
    // Each peer block is packed locally in the order of its destination rows (workspace: `aux`, aux_bytes()) and
    // deposited at its final address in the peer's `out` by ONE strided 3-D copy on a copy engine; the self block
    // is a plain local kernel.  Same barriers, same result as the direct-store kernel, but the link runs at the
    // copy engines' 735-755 GB/s instead of the ~680 GB/s of SM stores and keeps its rate while SM kernels use
    // the HBM, so the local transposition next to it can run beside it.  The pieces, for Plan::run_transpose_pair:
    //     dma_begin(out, s)            map the members' `out`, "every member's `out` is free" barrier
    //     dma_send(in, out, aux, p, q, s)  pack slice q of block p (stream s), then its copy on a copy stream
    //     dma_self(in, out, s)         my own block
    //     dma_signal(p, q) / dma_wait(r, q, s)  per-pair "slice has landed" flags instead of the group barrier
    //     dma_end(s, barrier)          join the copy streams (+ "all landed" barrier)
    bool dma_mode() const { return created_ && has_exchange_ && backend_ == BACKEND_NVLINK_FUSED && dma_; }
    bool dma_standalone() const { return dma_mode() && dma_standalone_; }
    int dma_begin(void* out, cudaStream_t stream);
    int dma_send(const void* in, void* out, void* aux, int peer, int sub, cudaStream_t stream);
    // slices of the block I send to member `peer` / receive from member `source` (geometry.h: dma_nsub)
    int n_subs_to(int peer) const { return peer == me_ ? 1 : (int)dma_subs_[(size_t)peer].size(); }
    int n_subs_from(int source) const {
        return source == me_ ? 1 : dma_nsub(send_by_member_[(size_t)source], recv_by_member_[(size_t)me_], es_, (int)members_.size());
    }
    int dma_self(const void* in, void* out, cudaStream_t stream);
    int dma_advance_signals(cudaStream_t stream) { return ctx_.peers->advance_epoch(members_, 6 + (comm_id_ - 1), stream); }
    int dma_signal(int peer, int sub);
    int dma_wait(int source, int sub, cudaStream_t stream);
    int dma_end(cudaStream_t stream, bool landed_barrier);
    int n_members() const { return (int)members_.size(); }
    int my_index() const { return me_; }
    const std::vector<Pencil>& recv_by_member() const { return recv_by_member_; }
    // Pieces of a LOCAL transposition cut by the members of the exchange next to it (geometry.h: local_box_for_peer):
    // piece `peer` writes (side 0, the exchange follows) or reads (side 1, the exchange precedes) exactly the elements
    // that travel between me and `peer` in that exchange.
    // (x_src -> x_dst) names the block of the exchange: side 0: my pencil after the local transposition -> the peer's
    // destination; side 1: the sender's source -> my pencil before the local transposition; slice `sub` of `nsub`.
    int local_piece(const void* in, void* out, int side, const Pencil& x_src, const Pencil& x_dst, int peer, int sub, int nsub,
                    cudaStream_t stream);

private:
    int execute_fused(void* in, void* out, cudaStream_t stream);
    std::vector<Pencil> send_by_member_;

    HandleContext ctx_;
    bool created_ = false, has_exchange_ = false, is_transpose_ = true;
    int backend_ = BACKEND_NCCL, comm_id_ = 1, me_ = 0;
    int64_t es_ = 0, aux_bytes_ = 0;
    long long send_elems_ = 0, remote_elems_ = 0, recv_elems_ = 0;
    int launches_ = 0;
    std::vector<int> members_;
    HandleGeometry geo_;
    std::unique_ptr<Kernel> pack_, unpack_;
    std::unique_ptr<NcclBackend> nccl_;
    // fused path
    std::vector<Box> fused_boxes_;  // per member, out_off relative to the member's `out`
    Family fused_family_ = FAM_NONE;
    std::map<const void*, std::unique_ptr<Kernel>> fused_;  // keyed by my `out` pointer
    long long evictions_ = 0;
    static constexpr size_t kMaxCachedDestinations = 256;
    Pencil send_;
    std::vector<Pencil> recv_by_member_;
    // chunk kernels keyed by (my `out` pointer, nchunks)
    std::map<std::pair<const void*, int>, std::vector<std::unique_ptr<Kernel>>> fused_chunks_;
    // Where every member receives: the members' `out` buffers mapped into this process.  The first call with
    // a given `out` (and the first after that address was re-allocated) is COLLECTIVE over the world
    // communicator (PeerRegistry::publish): like every dtFFT call it must be made by all ranks alike.
    // DTFFTB_ERROR_NOT_REGISTERED (on every rank alike) if some rank's buffer cannot be shared over cudaIpc.
    int peer_bases(void* out, std::vector<void*>* bases);
    struct PeerMap {
        std::vector<void*> bases;   // per member of the 1-D communicator
        std::vector<void*> opened;  // per world rank, for release()
        unsigned long long id = 0;  // buffer_id(out) when it was published
    };
    std::map<const void*, PeerMap> maps_;
    // copy-engine form
    bool dma_ = false;             // the copy-engine form is set up: pair pipelines use it
    bool dma_standalone_ = false;  // ... and so does a lone execute() (DTFFTB_FUSED_MODE=dma)
    struct DmaSub {
        DmaBlock blk;
        int pack_index = -1;  // box of dma_pack_ (and event) of this slice; -1 = empty
    };
    std::vector<std::vector<DmaSub>> dma_subs_;  // [member][slice]
    int copy_stream_of(int peer) const {
        const int P = (int)members_.size();
        return ((peer - me_ + P) % P + n_copy_streams_ - 1) % n_copy_streams_;
    }
    std::unique_ptr<Kernel> dma_pack_;      // in -> staging, one launch per peer
    std::unique_ptr<Kernel> dma_self_;      // my own block, in -> out
    static constexpr int kCopyStreams = 8;
    int n_copy_streams_ = 2;  // DTFFTB_DMA_STREAMS (1..4): peers are dealt round over this many copy streams
    cudaStream_t copy_streams_[kCopyStreams] = {};
    std::vector<cudaEvent_t> pack_done_;    // per slice
    cudaEvent_t copies_done_[kCopyStreams] = {};
    bool copy_used_[kCopyStreams] = {};
    const std::vector<void*>* dma_bases_ = nullptr;  // of the `out` of the exchange in flight (dma_begin)
    int execute_dma(void* in, void* out, cudaStream_t stream, void* aux);
    int ensure_dma_resources();
    // local pieces cut by the members of a neighbouring exchange: [side] keyed by member
    std::map<std::pair<int, int>, std::unique_ptr<Kernel>> peer_pieces_[2];
};

}  // namespace dtfftb
