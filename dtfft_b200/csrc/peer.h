// Peer-memory registry and device-side barrier for the NVLink direct-store backend.
//
// The reference has no counterpart on the NCCL path (its closest relatives are the
// NVSHMEM symmetric heap of backend_cufftmp, src/dtfft_backend_cufftmp.F90:39-120, and the
// MPI RMA windows of src/dtfft_backend_mpi.F90:397-431).  On one NVSwitch box every GPU
// can store straight into every peer's HBM, so a transposition can be ONE kernel that
// reads the local pencil and writes each element at its final place in the owning peer.
//
// * Buffers are "symmetric": every rank registers its buffers collectively and in the
//   same order (dtfft_mem_alloc does it automatically), so (slot, byte offset) names the
//   same logical buffer on every rank.  Handles travel as cudaIpcMemHandle_t through the
//   host allgather; each peer maps them once.
// * Ordering between GPUs uses a flags array per rank: a one-block kernel stores the
//   current epoch into every group member's flags with st.release.sys and spins on its
//   own flags with ld.acquire.sys.  Two such barriers bracket the remote stores
//   ("destination is free" / "data has landed").
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <vector>

#include "blocks.h"
#include "comm.h"

namespace dtfftb {

// Identity of the allocation behind a device pointer (0 if unknown): changes when the address is re-allocated.
unsigned long long buffer_id(const void* ptr);

class PeerRegistry {
public:
    ~PeerRegistry() { destroy(); }
    // Collective over `world`.  After it, available() tells whether every rank sits on
    // the same host with peer access to every other rank's device.
    int init(const Comm& world);
    bool available() const { return available_; }
    bool shared_device() const { return shared_device_; }  // test mode, see init()
    const char* why_unavailable() const { return why_; }

    // Collective, same order on every rank.  `ptr` may point inside a larger cudaMalloc
    // allocation (e.g. a torch caching-allocator segment).
    int register_buffer(void* ptr, size_t bytes, int* slot_out = nullptr);
    int unregister_buffer(void* ptr);  // collective
    bool resolve(const void* ptr, size_t bytes, int* slot, size_t* offset) const;
    // Collective over `world`, ANY device buffer (the drop-in case: the reference accepts every device
    // pointer, src/dtfft_plan.F90:1769-1795).  Every rank publishes the buffer it is about to receive into;
    // on return (*mapped)[r] is rank r's buffer in THIS process' address space (mine = ptr) and `opened`
    // holds what release() must close.  *ok == false (on every rank alike) when some rank could not export
    // its buffer through cudaIpc (stream-ordered or virtual-memory allocations) or could not map a peer's:
    // nothing stays mapped then.  An allocation is opened once per process however often it is published.
    int publish(void* ptr, size_t bytes, std::vector<void*>* mapped, std::vector<void*>* opened, bool* ok);
    void release(std::vector<void*>* opened);
    // Address of (slot, offset) of world rank `r` in THIS process' address space.
    void* peer_ptr(int r, int slot, size_t offset) const;

    // Enqueue a barrier among `members` (world ranks, must contain me) on `stream`.
    // `channel` separates independent groups (1-D communicator id 1..3, x2 phases).
    int barrier(const std::vector<int>& members, int channel, cudaStream_t stream);
    // Pairwise "landed" flags (copy-engine exchange with a consumer right behind it): one epoch counter per
    // (members, channel) in device memory, advanced once per exchange by advance_epoch(); signal() stores the
    // current epoch into flag [channel][me] of ONE member (enqueue it behind the copy that carries the data),
    // wait() spins until flag [channel][source] of mine has reached the current epoch.  Every member runs the same
    // sequence of exchanges, so the epochs agree; all three are graph-replayable like barrier().
    int advance_epoch(const std::vector<int>& members, int channel, cudaStream_t stream);
    // `slice` (0..kMaxSlices-1): the flag value is epoch * kMaxSlices + slice + 1, so the slices of one block, which
    // land in order, need only one flag.
    int signal(const std::vector<int>& members, int channel, int member_index, int slice, cudaStream_t stream);
    int wait(const std::vector<int>& members, int channel, int member_index, int slice, cudaStream_t stream);
    static constexpr int kMaxSlices = 16;
    // Collective.  Forget every barrier group and zero the flags: must be called when the 1-D
    // communicators change (process-grid search), because a new group starts at epoch 0 while the
    // flag rows of its channel may still hold the last epoch of a differently composed group.
    int reset_barriers();
    // Non-zero if a barrier ever timed out (peer missing or late); sticky.  Reads a word in mapped host
    // memory that the timed-out kernel wrote: no synchronisation, cheap enough for every API call.
    int error_state() const;
    // Device address of the sticky error word; the fused kernels poll it and store nothing once it is set.
    const unsigned long long* abort_flag() const {
        return flags_ ? reinterpret_cast<const unsigned long long*>(flags_ + (size_t)kChannels * world_.size()) : nullptr;
    }
    long long timeout_cycles() const { return timeout_cycles_; }
    void destroy();

    static constexpr int kChannels = 10;

private:
    struct Slot {
        void* local = nullptr;
        size_t bytes = 0;
        std::vector<void*> mapped;  // per world rank (mine = local)
        std::vector<void*> opened;  // IPC bases to close, per world rank
        bool live = false;
    };
    struct Group {
        uint64_t** d_peer_flags = nullptr;  // [n] device array: flags base of each member
        int* d_members = nullptr;           // [n] world ranks
        int n = 0;
        uint64_t* d_epoch = nullptr;        // device-side epoch counter of this group (graph-replayable)
    };
    int open_all(const void* base, size_t offset, Slot& s);
    Group* group_for(const std::vector<int>& members, int channel, int* rc);

    Comm world_;
    bool inited_ = false, available_ = false, shared_device_ = false;
    const char* why_ = "not initialised";
    std::vector<Slot> slots_;
    uint64_t* flags_ = nullptr;  // [kChannels][world] + error word
    uint64_t* h_err_ = nullptr;  // mirror of the error word in mapped pinned host memory
    uint64_t* d_err_host_ = nullptr;  // device address of h_err_
    long long timeout_cycles_ = 40ll * 1000 * 1000 * 1000;  // DTFFTB_PEER_TIMEOUT_MS (default 20 s) x SM clock
    int flags_slot_ = -1;
    std::map<std::pair<int, std::vector<int>>, Group> groups_;
};

}  // namespace dtfftb
