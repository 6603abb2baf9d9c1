// The plan object behind the public C API: the B200 replacement of dtfft_plan_t and its
// containers transpose_plan / reshape_plan (src/dtfft_plan.F90, src/dtfft_transpose_plan.F90,
// src/dtfft_reshape_plan.F90, src/dtfft_reshape_plan_base.F90).  Host logic only; all data
// movement happens in ReshapeHandle / Kernel / NcclBackend and in cuFFT.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <cufftXt.h>
#include <nccl.h>

#include <array>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "comm.h"
#include "fft_executor.h"
#include "geometry.h"
#include "handle.h"
#include "peer.h"

namespace dtfftb {

// dtfft_config_t (include/dtfft.h:1159-1389) with the defaults of src/dtfft_config.F90:644-669,
// except `platform`, which is CUDA here.
struct Config {
    bool enable_log = false, enable_z_slab = true, enable_y_slab = false;
    int32_t n_measure_warmup_iters = 2, n_measure_iters = 5;
    int platform = 2;
    void* stream = nullptr;
    int backend = BACKEND_NONE, reshape_backend = BACKEND_NONE;
    bool enable_datatype_backend = true, enable_mpi_backends = false, enable_pipelined_backends = true;
    bool enable_rma_backends = true, enable_fused_backends = true, enable_nccl_backends = true;
    bool enable_nvshmem_backends = true, enable_kernel_autotune = false, enable_fourier_reshape = false;
    int transpose_mode = 15, access_mode = -1;
};
Config& global_config();
// Struct values overridden by DTFFT_* environment variables (src/dtfft_config.F90:388-483, 788-802).
Config effective_config();

enum PlanKind : int { PLAN_C2C = 0, PLAN_R2C = 1, PLAN_R2R = 2 };

class Plan {
public:
    Plan() = default;
    ~Plan() { destroy(); }
    int create(PlanKind kind, int ndims, const int32_t* dims, const dtfft_pencil_t* pencil, const int* r2r_kinds,
               const dtfftb_comm_t* comm, int precision, int effort, int executor, bool dry = false);
    int execute(void* in, void* out, int execute_type, void* aux);
    int transpose(void* in, void* out, int ttype, void* aux);
    int reshape(void* in, void* out, int rtype, void* aux);
    int destroy();

    int get_local_sizes(int32_t* in_starts, int32_t* in_counts, int32_t* out_starts, int32_t* out_counts,
                        size_t* alloc_size) const;
    size_t alloc_size() const;
    size_t element_size() const;
    size_t alloc_bytes() const { return alloc_size() * element_size(); }
    size_t aux_bytes_transpose() const;
    size_t aux_bytes_reshape() const;
    size_t aux_bytes() const;
    int get_pencil(int layout, dtfft_pencil_t* p) const;
    int mem_alloc(size_t bytes, void** ptr);
    int mem_free(void* ptr);
    int register_buffer(void* ptr, size_t bytes);
    int unregister_buffer(void* ptr);
    int report() const;

    bool created() const { return created_; }
    int ndims() const { return ndims_; }
    const int32_t* dims() const { return user_dims_; }
    const int32_t* grid_dims() const { return comm_dims_; }
    bool z_slab() const { return is_z_slab_; }
    bool y_slab() const { return is_y_slab_; }
    bool reshape_enabled() const { return is_reshape_enabled_; }
    int executor() const { return executor_; }
    int precision() const { return precision_; }
    int backend() const { return backend_; }
    int reshape_backend() const { return reshape_backend_; }
    cudaStream_t stream() const { return stream_; }
    PlanKind kind() const { return kind_; }
    void last_stats(int64_t* launches, int64_t* local_bytes, int64_t* remote_bytes) const {
        *launches = stat_launches_, *local_bytes = stat_local_, *remote_bytes = stat_remote_;
    }
    int peer_error() { return peers_.error_state(); }
    // Host-only test hook (dry plans): switch a default 3-D decomposition to the grid 1 x g1 x g2
    // the way the grid search does.
    int dry_set_grid(int g1, int g2);
    // Stage overlap of the cuFFT executor with the NVLINK_FUSED exchange: number of chunks
    // (<= 1 disables) and CTAs of the persistent exchange kernel (0 = one per SM).
    void set_overlap(int nchunks, int ctas) {
        overlap_chunks_ = nchunks, overlap_ctas_ = ctas, overlap_user_set_ = true;
        drop_graphs();
    }
    // dtfft_execute replays a CUDA graph captured on the second call with the same buffers
    // (single-GPU and NVLINK_FUSED plans; env DTFFTB_GRAPHS=0 disables).
    void set_graphs(bool on) {
        graphs_enabled_ = on;
        drop_graphs();
    }
    int64_t graph_replays() const { return stat_graph_replays_; }
    int64_t fallbacks() const { return stat_fallbacks_; }
    // How one transposition moves its data: 0 local kernel only, 1 NCCL, 2 direct-store kernel, 3 copy engines,
    // 4 direct-store kernel on its own, copy engines when pipelined with the local transposition next to it
    // (*n_slices = copies this rank issues per execute).
    int exchange_form(int ttype, int* form, int* n_slices) const {
        auto it = handles_.find(ttype);
        if (it == handles_.end()) return DTFFT_ERROR_INVALID_TRANSPOSE_TYPE;
        const ReshapeHandle& h = *it->second;
        *n_slices = 0;
        if (!h.has_exchange()) *form = 0;
        else if (h.backend() != BACKEND_NVLINK_FUSED) *form = 1;
        else if (!h.dma_mode()) *form = 2;
        else {
            *form = h.dma_standalone() ? 3 : 4;
            for (int i = 0; i < h.n_members(); ++i)
                if (i != h.my_index()) *n_slices += h.n_subs_to(i);
        }
        return DTFFT_SUCCESS;
    }
    int overlap_chunks() const { return overlap_chunks_; }
    int64_t overlapped_stages() const { return stat_overlapped_; }

    // Exchange geometry of one transposition (dtfft_transpose_t) or reshape (dtfft_reshape_t) on
    // this rank: the reference's neighbor_data tables (transposes) and the fused-path boxes.
    struct ExchangeDescription {
        std::vector<int> members;
        int me = 0;
        int64_t element_bytes = 0;
        HandleGeometry geo;
        std::vector<Box> fused;
        bool fused_transposing = false;
    };
    int describe_exchange(int type, ExchangeDescription* d) const;
    // Brick <-> pencil reshape over the NCCL backends: pack / unpack boxes, exchange tables and
    // the pack-free / unpack-free flags (geometry.h: reshape_geometry).
    int describe_reshape(int rtype, std::vector<int>* members, int* me, ReshapeGeometry* g) const;
    // Boxes of chunk k of n of the stage-overlapped fused transposition `ttype` (geometry.h: chunk_boxes).
    int describe_chunk(int ttype, int k, int nchunks, std::vector<int>* members, std::vector<Box>* boxes,
                       long long* chunk_offset) const;
    // Copy-engine form of a transposition and the pieces of the LOCAL transposition next to it, cut by the (peer,
    // slice) blocks of that exchange (geometry.h: DmaBlock, local_box_for_block): side 0 = producer, 1 = consumer.
    struct DmaEntry {
        int member = 0, sub = 0, nsub = 1;
        DmaBlock blk;
        Box fused;
    };
    int describe_dma(int ttype, std::vector<int>* members, int* me, std::vector<DmaEntry>* entries) const;
    int describe_peer_piece(int t_local, int t_exchange, int side, int peer, int sub, int* nsub_out, Box* box) const;
    std::vector<int> transpose_types() const;

private:
    struct UserPencil {  // X-aligned starts / counts of one rank, as supplied by the user
        int32_t starts[3], counts[3];
    };
    struct HandleSpec {
        int ttype = 0, rtype = 0, comm_id = 1, me = 0;
        int64_t es = 0;
        std::vector<int> members;
        std::vector<Pencil> send, recv;
    };
    int handle_spec(int type, HandleSpec* hs) const;
    int choose_decomposition(const dtfft_pencil_t* pencil);
    std::vector<int> group_members(int rank, int comm_id) const;
    int build_pencils();
    int init_nccl();
    int build_handles(int backend, std::map<int, std::unique_ptr<ReshapeHandle>>& into);
    int build_reshape_handles(int backend);
    int build_reshape_handles(int backend, std::map<int, std::unique_ptr<ReshapeHandle>>& into);
    int autotune_backend();
    int autotune_reshape_backend();  // DTFFT_EXHAUSTIVE, src/dtfft_reshape_plan.F90:206-222
    // DTFFT_MEASURE / DTFFT_PATIENT process-grid search (autotune_grid_decomposition,
    // src/dtfft_transpose_plan.F90:391-540); `all_backends` also times every enabled backend per grid.
    int autotune_grid(bool all_backends);
    void set_grid(int g1, int g2);
    std::vector<int> backend_candidates() const;
    int choose_overlap();
    int time_backend(int backend, double* ms, bool reshapes = false);
    int create_ffts();
    int check_aux(void* aux, bool from_execute, void** aux1, void** aux2);
    int check_device_ptrs(const void* a, const void* b, const void* c) const;
    int run_transpose(int ttype, void* in, void* out, void* aux);
    int run_reshape(int rtype, void* in, void* out, void* aux);
    int run_fft(int dim, void* a, void* b, int sign);
    // FFT a -> b followed by the transposition b -> c.  With the NVLINK_FUSED backend the two
    // are pipelined chunk by chunk over two streams (stage overlap); otherwise run back to back.
    int run_transpose_pair(int t1, void* a, void* b, int t2, void* c, void* aux);
    int ensure_overlap_resources(long long nch, bool consumer);
    int run_fft_transpose(int dim, void* a, void* b, int sign, int ttype, void* c, void* aux);
    int execute_schedule(void* in, void* out, bool fwd, void* a1, void* a2, bool inplace);
    bool graphs_usable() const;
    void drop_graphs();
    void forget_buffer_caches();
    int execute_2d(void* in, void* out, bool fwd, void* aux, void* aux2);
    int execute_z_slab(void* in, void* out, bool fwd, void* aux, bool inplace, void* aux2);
    int execute_generic(void* in, void* out, bool fwd, void* aux, void* aux2);
    int execute_2d_reshape(void* in, void* out, bool fwd, void* aux, void* aux2);
    int execute_z_slab_reshape(void* in, void* out, bool fwd, void* aux, void* aux2);
    int execute_generic_reshape(void* in, void* out, bool fwd, void* aux, void* aux2);
    void log(const char* fmt, ...) const;

    bool created_ = false, dry_ = false;
    PlanKind kind_ = PLAN_C2C;
    Config cfg_;
    Comm comm_;
    int ndims_ = 0;
    int32_t user_dims_[3] = {1, 1, 1};   // what the user asked for (real extents for R2C)
    int32_t dims_[3] = {1, 1, 1};        // extents the transposes work on (complex for R2C)
    int32_t comm_dims_[3] = {1, 1, 1};
    int precision_ = 1, effort_ = 0, executor_ = 0;
    int r2r_kinds_[3] = {-1, -1, -1};
    bool is_transpose_plan_ = true, is_z_slab_ = false, is_y_slab_ = false;
    bool is_reshape_enabled_ = false, is_final_reshape_enabled_ = false;
    int64_t base_storage_ = 16, base_storage_init_ = 16;
    int backend_ = BACKEND_NCCL, reshape_backend_ = BACKEND_NCCL;
    cudaStream_t stream_ = nullptr;
    bool own_stream_ = false;

    bool has_user_pencil_ = false;
    std::vector<std::array<int32_t, 3>> coords_;        // per world rank: coordinates in the pencil grid
    std::vector<std::array<int32_t, 3>> brick_coords_;  // per world rank: coordinates in the user's brick grid
    int32_t brick_grid_[3] = {1, 1, 1};
    std::vector<UserPencil> xpencil_from_bricks_;       // per world rank: X pencil chosen by from_bricks
    std::vector<UserPencil> user_pencils_;            // per world rank (bricks or X pencils)
    std::vector<std::array<Pencil, 3>> pencils_;       // per world rank: X, Y, Z pencils
    std::vector<Pencil> real_pencils_;                 // per world rank (R2C)
    std::vector<std::array<Pencil, 2>> bricks_;        // per world rank: X bricks, Z bricks (reshape)

    ncclComm_t nccl_ = nullptr;
    PeerRegistry peers_;
    std::map<int, std::unique_ptr<ReshapeHandle>> handles_;   // keyed by dtfft_transpose_t
    std::map<int, std::unique_ptr<ReshapeHandle>> rhandles_;  // keyed by dtfft_reshape_t
    // NCCL stand-ins of the NVLINK_FUSED handles, built on first need: a caller's buffer that cudaIpc cannot
    // export (stream-ordered or virtual-memory allocations) still has to work, like every device pointer
    // does in the reference (src/dtfft_plan.F90:1769-1795).  Every rank falls back together (publish agrees).
    std::map<int, std::unique_ptr<ReshapeHandle>> fb_handles_, fb_rhandles_;
    int fallback_execute(bool reshape, int type, void* in, void* out, void* aux);
    int64_t stat_fallbacks_ = 0;
    std::unique_ptr<FftExecutor> fft_[3];
    int fft_mapping_[3] = {0, 1, 2};

    struct Alloc {
        void* ptr;
        size_t bytes;
        bool nccl;       // ncclMemAlloc'ed
        void* reg;       // ncclCommRegister handle
        bool peer;       // registered with the peer registry
    };
    std::vector<Alloc> allocs_;
    void* aux_ptr_ = nullptr;
    bool is_aux_alloc_ = false;
    int64_t stat_launches_ = 0, stat_local_ = 0, stat_remote_ = 0, stat_overlapped_ = 0;
    int64_t stat_eager_only_ = 0;  // stages of the last execute whose two-stream pipeline must not be replayed from a graph
    // CUDA-graph replay of dtfft_execute
    struct GraphKey {
        const void *in, *out, *aux;
        bool fwd;
        bool operator<(const GraphKey& o) const {
            return std::tie(in, out, aux, fwd) < std::tie(o.in, o.out, o.aux, o.fwd);
        }
    };
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        bool failed = false;
        int64_t launches = 0, local = 0, remote = 0, overlapped = 0;
        long long evictions = 0;  // handle_evictions() when the graph was captured
        unsigned long long ids[3] = {0, 0, 0};  // buffer_id of (in, out, aux) at capture: a re-allocation voids the graph
    };
    long long handle_evictions() const;
    std::map<GraphKey, GraphEntry> graphs_;
    bool graphs_enabled_ = true;
    int64_t stat_graph_replays_ = 0;
    // stage overlap
    int overlap_chunks_ = 1, overlap_ctas_ = 0;
    // Peer-by-peer pipelining of a local transposition with a copy-engine exchange next to it (run_transpose_pair);
    // DTFFTB_PAIR_OVERLAP=0 runs the two transpositions one after the other
    bool pair_overlap_ = true;
    cudaStream_t bar_stream_ = nullptr;
    std::vector<cudaEvent_t> landed_events_;
    cudaEvent_t pair_start_ = nullptr;
    bool overlap_user_set_ = false;
    cudaStream_t xfer_stream_ = nullptr;
    std::vector<cudaEvent_t> chunk_events_;
    cudaEvent_t xfer_done_ = nullptr;
};

}  // namespace dtfftb
