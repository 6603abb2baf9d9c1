// Host-side kernel object: the B200 replacement of the reference's kernel_device
// (src/dtfft_kernel_device.F90) behind abstract_kernel's contract
// (src/dtfft_abstract_kernel.F90:219-403).  Plain C++ so that the plan layer can use it
// directly; the C ABI in kernel_api.cu is a thin shell over this class.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "blocks.h"
#include "geometry.h"  // Box
#include "kernels.cuh"

namespace dtfftb {

enum KernelType : int {
    K_DUMMY = -1,
    K_PACK = 1,
    K_COPY_PIPELINED = 2,
    K_UNPACK = 3,
    K_COPY = 4,
    K_UNPACK_PIPELINED = 5,
    K_PACK_PIPELINED = 6,
    K_PERMUTE_FORWARD = 7,
    K_PERMUTE_BACKWARD = 8,
    K_PERMUTE_BACKWARD_START = 9,
    K_PERMUTE_BACKWARD_END = 10,
    K_PERMUTE_BACKWARD_END_PIPELINED = 11,
    K_PACK_FORWARD = 12,
    K_PACK_BACKWARD = 13,
    K_UNPACK_FORWARD = 15,
    K_UNPACK_FORWARD_PIPELINED = 16,
    K_UNPACK_BACKWARD = 17,
    K_UNPACK_BACKWARD_PIPELINED = 18,
};

enum Family : int { FAM_NONE = 0, FAM_COPY = 1, FAM_T = 2, FAM_R = 3 };

bool is_per_neighbor_kind(int t);
bool needs_neighbor_data(int t);
Family family_of(int t);
int effective_type(int t, int ndims);  // 2-D backward -> forward remap
// Box of neighbour `nd5` (or the whole buffer when nd5 == nullptr) for kind `t`.
Box make_box(int t, int ndims, const int32_t* dims, const int32_t* nd5);

struct DeviceTable {
    long long offset = 0;  // index of the first BlockDesc in the kernel's device array
    int nblocks = 0;
    long long total_items = 0;
};

class Kernel {
public:
    Kernel() = default;
    ~Kernel();
    Kernel(const Kernel&) = delete;
    Kernel& operator=(const Kernel&) = delete;

    // Returns a dtfft error code / DTFFTB_ERROR_*.
    int create(int ndims, const int32_t* dims, int kernel_type, int64_t base_storage, const int32_t* neighbor_data,
               int n_neighbors, int effort, bool force_effort);
    // Kernel over explicit boxes (one per peer, empty boxes allowed): used by the plan layer
    // for the fused NVLink path and the brick reshapes, whose geometry is derived from
    // global-index intersections rather than from neighbor_data.
    int create_boxes(Family family, int64_t base_storage, const std::vector<Box>& boxes);
    int execute(const void* in, void* out, cudaStream_t stream, int neighbor, bool sync);
    int execute_all(const void* in, void* out, cudaStream_t stream);
    int set_peer_out(void* const* out_bases, const int64_t* out_displs_override);
    int set_tile(int ka, int kb, int rows);
    // Run as a persistent kernel of at most `max_ctas` CTAs (0 = no limit).  Used when the launch
    // shares the GPU with another stream (stage overlap): the NVLink-bound stores need few SMs.
    void set_grid_limit(int max_ctas) { grid_limit_ = max_ctas; }
    // Fused NVLink kernels: device address of the group's sticky peer-error word (peer.h).  Takes effect at
    // the next table rebuild (set_peer_out / set_tile); once the word is set the kernel stores nothing.
    void set_abort_flag(const unsigned long long* flag) { abort_flag_ = flag; }
    int sm_count() const { return sm_count_; }
    int autotune(const void* in, void* out, cudaStream_t stream, int n_warmup, int n_iters, float* best_ms);
    // One entry per tile candidate the last autotune() timed (kernel_device.F90:385-389 logs the same pair).
    struct AutotuneEntry {
        TileCfg cfg;
        float ms;
        double gbs;  // 2 x bytes moved / time
    };
    void set_autotune_log(std::vector<AutotuneEntry>* log) { autotune_log_ = log; }
    void destroy();

    // Host-only mode (tests on CPU boxes): geometry and tables are built exactly as for a launch but
    // nothing touches a device; execute() fails.  Set BEFORE create / create_boxes.
    void set_dry(bool d) { dry_ = d; }
    bool dry() const { return dry_; }
    // The table a launch would hand to the device (dry kernels only): family R `unit` in {4, 8, 16}
    // (ignored for family T), `neighbor` 1-based or 0 for the all-peer table.  Returns nullptr when
    // the table does not exist (unit wider than the geometry allows).
    const BlockDesc* host_table(int unit, int neighbor, DeviceTable* t, int launch[3]) const;

    bool is_noop() const { return noop_; }
    Family family() const { return family_; }
    int type() const { return type_; }
    int n_neighbors() const { return P_; }
    int64_t element_size() const { return es_; }
    // floats (4-byte units) moved for neighbour n (1-based), abstract_kernel.F90:388-397
    long long csize(int neighbor) const;
    void get_info(int* family, int* unit, int* tile_a, int* tile_b, int* threads, int64_t* n_items) const;
    long long bytes_moved() const;  // algorithmic payload bytes of an all-peer launch (one direction)

private:
    int rebuild_tables();
    int launch(const DeviceTable& t, int unit, const void* in, void* out, cudaStream_t stream);
    int pick_unit(const void* in, const void* out) const;

    bool created_ = false, noop_ = true, custom_ = false, dry_ = false;
    std::vector<BlockDesc> host_tables_;  // dry mode: what would have been uploaded
    int ndims_ = 0, type_ = K_DUMMY, P_ = 0;
    int32_t dims_[3] = {1, 1, 1};
    int64_t es_ = 0;
    Family family_ = FAM_NONE;
    std::vector<int32_t> nd_;     // P x 5, row per neighbour
    std::vector<Box> boxes_;      // per neighbour (or one)
    std::vector<void*> peer_out_; // optional per-neighbour out base
    std::vector<long long> peer_out_displ_;
    TileCfg tile_{1, 1, 8};
    std::vector<AutotuneEntry>* autotune_log_ = nullptr;
    const unsigned long long* abort_flag_ = nullptr;
    bool align_b_ = true;  // DTFFTB_ALIGN_TILES=0: tile grid anchored at the box origin (A/B of BlockDesc::bshift)
    int tx_ = 32;
    int tx_slot_[3] = {32, 32, 32};
    int unit_geo_ = 4;  // widest unit the geometry allows (family R)
    int grid_cap_ = 148 * 8;
    int grid_limit_ = 0;
    int sm_count_ = 148;
    // device tables: index 0 -> family T (unit = element) or family R unit 4; 1 -> unit 8; 2 -> unit 16
    BlockDesc* d_blocks_ = nullptr;
    DeviceTable all_[3];
    std::vector<DeviceTable> single_[3];
};

}  // namespace dtfftb
