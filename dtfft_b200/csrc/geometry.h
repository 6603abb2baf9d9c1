// Decomposition and per-peer block geometry (CUDA-free host logic).
//
// Restates the layout math the kernels are parameterised by:
//   get_local_size            src/dtfft_pencil.F90:237-279
//   pencil permutations        src/dtfft_transpose_plan.F90:1046-1082
//   default grid choice        src/dtfft_transpose_plan.F90:170-203
//   create_pencils_and_comm    src/dtfft_transpose_plan.F90:1084-1131
//   1-D communicator of a transposition   src/dtfft_abstract_reshape_handle.F90:156-188
//   neighbor_data / counts / displs / kernel kinds   src/dtfft_reshape_handle_generic.F90:291-640
// MPI is replaced by a world allgather callback (comm.h); sub-communicators are index
// lists into the world.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

namespace dtfftb {

enum TransposeType : int { T_X_TO_Y = 1, T_Y_TO_X = -1, T_Y_TO_Z = 2, T_Z_TO_Y = -2, T_X_TO_Z = 3, T_Z_TO_X = -3 };
enum ReshapeType : int { R_X_BRICKS_TO_PENCILS = 11, R_X_PENCILS_TO_BRICKS = 12, R_Z_PENCILS_TO_BRICKS = 13, R_Z_BRICKS_TO_PENCILS = 14 };

constexpr int kDefTileSize = 32;  // DEF_TILE_SIZE, src/dtfft_parameters.F90 (Z-slab rule on CUDA)

struct Pencil {
    int aligned_dim = 0;  // 1-based like the reference
    int ndims = 0;
    int32_t starts[3] = {0, 0, 0};
    int32_t counts[3] = {1, 1, 1};
    bool is_even = true;
    bool is_distributed = false;
    long long size() const {
        long long s = 1;
        for (int i = 0; i < ndims; ++i) s *= counts[i];
        return s;
    }
};

// One peer's box in ELEMENTS (a = input-contiguous axis).
struct Box {
    long long in_off = 0, out_off = 0;
    long long n0 = 0, n1 = 1, n2 = 1;
    long long is1 = 0, is2 = 0;
    long long os0 = 1, os1 = 0, os2 = 0;
    bool empty() const { return n0 <= 0 || n1 <= 0 || n2 <= 0; }
    long long volume() const { return n0 * n1 * n2; }
};

void local_size(int n_global, int comm_dim, int comm_rank, int32_t* start, int32_t* count);
// MPI_Dims_create for the shapes dtFFT asks for (zeros are free entries).
void dims_create(int nnodes, int ndims, int32_t* dims);

struct GridChoice {
    int32_t comm_dims[3] = {1, 1, 1};
    bool is_z_slab = false, is_y_slab = false;
    bool invalid_grid = false;
};
GridChoice choose_grid(int ndims, const int32_t* dims, int comm_size, bool cuda, bool z_slab_enabled,
                       bool y_slab_enabled);

// Process grids 1 x g1 x g2 tried by the DTFFT_MEASURE / DTFFT_PATIENT grid search, in the
// reference's order (autotune_grid_decomposition, src/dtfft_transpose_plan.F90:456-500: for every
// divisor i <= sqrt(P): (i, P/i) then (P/i, i)), minus the grids autotune_grid rejects because a
// pencil would have fewer points than ranks along a split axis (:600-607).
std::vector<std::pair<int, int>> grid_candidates(const int32_t* dims, int comm_size);

void cart_coords(int rank, int ndims, const int32_t* comm_dims, int32_t* coords);
int cart_rank(int ndims, const int32_t* comm_dims, const int32_t* coords);
// Global ranks of 1-D communicator `comm_id` (1 = whole grid, d >= 2 = grid dimension d) containing `rank`.
std::vector<int> comm_members(int rank, int ndims, const int32_t* comm_dims, int comm_id);
// dperm[d][j]: global axis of local axis j of pencil d; cperm[d][j]: grid dimension splitting it (0-based).
void permutations(int ndims, int dperm[3][3], int cperm[3][3]);
// The ndims pencils of `rank` for the default (no user pencil) decomposition.
void make_pencils(int ndims, const int32_t* dims, const int32_t* comm_dims, int rank, Pencil out[3]);

int transpose_comm_id(int ttype);
void transpose_pencil_ids(int ttype, int* send, int* recv);  // 0-based pencil indices

struct HandleGeometry {
    int ttype = 0;            // dtfft_transpose_t value, 0 for reshapes
    int rtype = 0;            // dtfft_reshape_t value, 0 for transposes
    int ndims = 0;
    int comm_size = 1, comm_rank = 0;
    std::vector<int> members;  // global ranks of the 1-D communicator
    int32_t send_dims[3] = {1, 1, 1}, recv_dims[3] = {1, 1, 1};
    int pack_kernel = -1, unpack_kernel = -1;  // kernel_type_t values (-1 = dummy / none)
    std::vector<int32_t> send_nd, recv_nd;     // 5 x P, row per peer
    std::vector<int64_t> send_counts, send_displs, recv_counts, recv_displs;  // elements
    bool has_exchange = false, is_pipelined = false, is_fused = false;
    bool is_pack_free = false, is_unpack_free = false;
    int reshape_strat = 0;
};

// Geometry of transposition `ttype` for member `me` of a 1-D communicator whose members'
// send / recv pencils are given in sub-communicator order.
HandleGeometry transpose_geometry(int ttype, const std::vector<Pencil>& send_by_member,
                                  const std::vector<Pencil>& recv_by_member, int me, const std::vector<int>& members,
                                  bool pipelined, bool fused);

// ---- geometry from global-index intersections (fused NVLink path, brick reshapes) ----
// A rank's local array: local axis j is global axis `axis[j]`, holding global indices
// [starts[j], starts[j] + counts[j]); column-major with local axis 0 fastest.
struct RankLayout {
    int ndims = 0;
    int axis[3] = {0, 1, 2};
    int32_t starts[3] = {0, 0, 0};
    int32_t counts[3] = {1, 1, 1};
    long long volume() const {
        long long v = 1;
        for (int j = 0; j < ndims; ++j) v *= counts[j];
        return v;
    }
};
RankLayout layout_of(const Pencil& p);
// The part of `src` that lands in `dst`, as a strided box: element (global g) is read at
// in_off + sum_j (g - src.start) * src.stride and written at out_off + ... of dst.
// *transposing = the fastest axes differ (family T), else family R.
Box intersect_box(const RankLayout& src, const RankLayout& dst, bool* transposing);
// Layout of the contiguous slot that carries the (src -> dst) block between two ranks:
// the intersection, stored in the axis order of `order_like`.
RankLayout slot_layout(const RankLayout& src, const RankLayout& dst, const RankLayout& order_like);

// Stage overlap of the fused path: chunk k of n along the SLOWEST axis of the source pencil, as a
// transposition of its own.  boxes[i] = part of the chunk owned by member i afterwards, with in_off
// relative to the chunk's first element (returned in *chunk_offset, elements of the full pencil).
std::vector<Box> chunk_boxes(const Pencil& send, const std::vector<Pencil>& recv_by_member, int k, int nchunks,
                             long long* chunk_offset);

// ---- copy-engine form of a direct-store block (handle.h, DMA mode) ------------------------------
// On B200 a kernel's remote stores leave as 128-byte NVLink writes and top out near 690 GB/s per direction,
// and they collapse (to ~420 GB/s) when an HBM-bound kernel runs beside them; a copy engine moves the same
// rows -- even 1 KB rows at an 8 KB pitch -- at 735-755 GB/s and keeps ~640 GB/s under that load
// (profiles/r02c_nvlink_probe_n2.md).  So a peer block can travel as: a LOCAL pack of the block in the order
// of its destination rows (one pass through HBM), then ONE strided 3-D copy that deposits every row at its
// final address in the peer's array (no unpack).  `pack` = the block's source box with packed destination
// strides (relative to the staging buffer, block at `staging_off`); the copy moves `planes` x `rows` rows of
// `run` elements from the staging block (dense) to dst_off + row * dst_pitch + plane * dst_plane_rows *
// dst_pitch.  ok == false: the block's two outer destination strides are not multiples of one another, it
// cannot be one pitched 3-D copy.
struct DmaBlock {
    Box pack;
    long long run = 0, rows = 0, planes = 0;
    long long dst_off = 0, dst_pitch = 0, dst_plane_rows = 0;
    bool ok = false;
};
DmaBlock dma_block(const Box& b, bool transposing, long long staging_off);

// A peer block is cut into `nsub` slices along the SLOWEST axis of the sender's source pencil so that the pack of
// slice s + 1 runs beside the copy of slice s (with few peers a whole block would serialise pack and copy).
// Sender and receiver derive the same count from what both know: the block's bytes and its extent along that axis
// (a slice keeps at least 32 elements of it, a full tile width wherever that axis is somebody's fastest one).
// About 8 copies per rank and exchange (8 / (group size - 1) slices per block), none below DTFFTB_DMA_SUB_BYTES
// (default 4 MiB); setting that variable lifts the per-group rule (at most 8 slices per block then).
int dma_nsub(const Pencil& sender_src, const Pencil& receiver_dst, int64_t base_storage, int group_size);
// Global index range [lo, hi) of slice s of n of the block (sender_src -> receiver_dst) along that axis (*axis).
void dma_sub_range(const Pencil& sender_src, const Pencil& receiver_dst, int s, int nsub, int* axis, long long* lo,
                   long long* hi);
// The block (or its slice) as a direct-store box: intersect_box clipped to the slice.
Box block_box(const Pencil& sender_src, const Pencil& receiver_dst, int s, int nsub, bool* transposing);

// The part of a LOCAL transposition `send -> recv` (one rank in its communicator) that writes exactly what
// member `peer` of the exchanging transposition `recv -> next_by_member[...]` will be sent: the global-index
// range of `next_by_member[peer]` clipped out of the local transposition.  Lets the local transposition run
// peer by peer in front of the packs of a DMA-mode exchange (Plan::run_transpose_pair).
Box local_box_for_peer(const Pencil& send, const Pencil& recv, const Pencil& next_of_peer);
// Same, restricted to slice s of n of the block (x_src -> x_dst) of the exchange: with side 0 the exchange follows
// (x_src = my pencil after the local transposition, x_dst = the peer's destination), with side 1 it precedes
// (x_src = the sender's source, x_dst = my pencil before the local transposition).
Box local_box_for_block(const Pencil& send, const Pencil& recv, const Pencil& x_src, const Pencil& x_dst, int s, int nsub);

// ---- brick <-> pencil reshape over NCCL: pack -> all-to-all(v) -> unpack -------------------
// Block (me -> i) is the global-index intersection of my source with i's destination, carried
// in a contiguous slot in DESTINATION axis order (both sides of a reshape share the axis order,
// so this is the reference's wire format: box (n1,n2,n3) packed densely, :343-376, 536-567).
//   pack_boxes[i]   : my source -> slot i at send_displs[i]     (kernel `pack`)
//   unpack_boxes[i] : slot i at recv_displs[i] -> my destination (kernel `unpack`)
// is_pack_free / is_unpack_free restate src/dtfft_reshape_handle_generic.F90:261-266, 479-484:
// the reference's predicate must hold on EVERY member (its allreduce) -- evaluated here from the
// gathered layouts -- and, as a safety net the reference does not need, every member's boxes
// must really be identity placements (block i already lies at its exchange displacement).
// reshape_strat: 1 = z split, 2 = y split, 3 = 2-D split (:267-289).
struct ReshapeGeometry {
    std::vector<Box> pack_boxes, unpack_boxes;
    std::vector<int64_t> send_counts, send_displs, recv_counts, recv_displs;  // elements
    bool is_pack_free = false, is_unpack_free = false;
    int reshape_strat = 0;
};
ReshapeGeometry reshape_geometry(int rtype, const std::vector<Pencil>& send_by_member,
                                 const std::vector<Pencil>& recv_by_member, int me);

}  // namespace dtfftb
