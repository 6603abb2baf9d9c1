/*
 * dtfft_b200.h -- C ABI of the B200-native dtFFT reshape path (libdtfft_b200.so).
 *
 * This is the drop-in boundary: plain C, pointers and sizes only.  Each group of entry
 * points replaces one Fortran class of the reference (ShatrovOA/dtFFT v3.2.0); the
 * reference interface it stands in for is cited as file:line.  The Fortran host code of
 * the reference reaches these through iso_c_binding (see INTEGRATION.md for the stubs).
 *
 * All functions return 0 (DTFFT_SUCCESS) or a dtfft_error_t value from
 * include/dtfft_config.h.in:82-151 of the reference; CUDA / NCCL failures, which the
 * reference turns into MPI_Abort (src/include/_dtfft_cuda.h:7-21), are returned as
 * DTFFTB_ERROR_CUDA_BASE - cudaError_t (resp. DTFFTB_ERROR_NCCL_BASE - ncclResult_t) so
 * that the caller decides.  Not thread-safe (like the reference): one call at a time.
 */
#ifndef DTFFT_B200_H
#define DTFFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTFFTB_ERROR_CUDA_BASE (-10000)
#define DTFFTB_ERROR_NCCL_BASE (-20000)
#define DTFFTB_ERROR_INTERNAL (-30000) /* invariant violated (reference: INTERNAL_ERROR) */
#define DTFFTB_ERROR_NOT_REGISTERED (-30001) /* NVLINK_FUSED backend: `out` cannot be shared with the peers through cudaIpc and no NCCL
                                              * communicator exists to stand in (plans always have one; plugin-level users may not) */
#define DTFFTB_ERROR_COMM (-30002) /* the host allgather callback failed */
#define DTFFTB_ERROR_PEER_TIMEOUT (-30003) /* NVLINK_FUSED backend: a group member did not reach a device barrier within
                                            * DTFFTB_PEER_TIMEOUT_MS (default 20000): the kernels of the group stored nothing
                                            * from then on; fatal for the plan, every later call returns this code */

/* ------------------------------------------------------------------------------------
 * Process group handed to the plan layer instead of an MPI_Comm.  The reference needs its
 * communicator only for metadata (Allgather of pencil extents, broadcast of the NCCL id,
 * MPI_Allreduce of timings: src/dtfft_reshape_handle_generic.F90:143-144,
 * src/dtfft_abstract_backend.F90:437-441, src/dtfft_reshape_plan_base.F90:588-706); all of
 * that is expressed with one collective.  `allgather` must copy `bytes` bytes from `send`
 * of every rank into `recv` (size * bytes, rank order) and return 0.  `cart_ndims` > 0
 * describes a user process grid (the MPI_Cart_create case of
 * src/dtfft_transpose_plan.F90:128-170), row-major rank order like MPI.
 * ---------------------------------------------------------------------------------- */
typedef struct dtfftb_comm_s {
    int32_t rank;
    int32_t size;
    void* ctx;
    int (*allgather)(void* ctx, const void* send, void* recv, int64_t bytes);
    int32_t cart_ndims;
    int32_t cart_dims[3];
} dtfftb_comm_t;

/* kernel_type_t values: src/dtfft_abstract_kernel.F90:59-98 */
enum {
    DTFFTB_KERNEL_DUMMY = -1,
    DTFFTB_KERNEL_PACK = 1,
    DTFFTB_KERNEL_COPY_PIPELINED = 2,
    DTFFTB_KERNEL_UNPACK = 3,
    DTFFTB_KERNEL_COPY = 4,
    DTFFTB_KERNEL_UNPACK_PIPELINED = 5,
    DTFFTB_KERNEL_PACK_PIPELINED = 6,
    DTFFTB_KERNEL_PERMUTE_FORWARD = 7,
    DTFFTB_KERNEL_PERMUTE_BACKWARD = 8,
    DTFFTB_KERNEL_PERMUTE_BACKWARD_START = 9,
    DTFFTB_KERNEL_PERMUTE_BACKWARD_END = 10,
    DTFFTB_KERNEL_PERMUTE_BACKWARD_END_PIPELINED = 11,
    DTFFTB_KERNEL_PACK_FORWARD = 12,
    DTFFTB_KERNEL_PACK_BACKWARD = 13,
    /* host-only in the reference (src/dtfft_kernel_device.F90:72-74); available on device here */
    DTFFTB_KERNEL_UNPACK_FORWARD = 15,
    DTFFTB_KERNEL_UNPACK_FORWARD_PIPELINED = 16,
    DTFFTB_KERNEL_UNPACK_BACKWARD = 17,
    DTFFTB_KERNEL_UNPACK_BACKWARD_PIPELINED = 18
};

/* ------------------------------------------------------------------------------------
 * Kernel plugin surface -- replaces kernel_device / nvrtc_module / nvrtc_block_optimizer
 * / nvrtc_module_cache (src/dtfft_kernel_device.F90:45-178, src/dtfft_nvrtc_module.F90,
 * src/dtfft_nvrtc_block_optimizer.F90, src/dtfft_nvrtc_module_cache.F90) behind the
 * deferred interface of abstract_kernel (src/dtfft_abstract_kernel.F90:137-163).
 * ---------------------------------------------------------------------------------- */
typedef struct dtfftb_kernel_s* dtfftb_kernel_t;

/* abstract_kernel%create (src/dtfft_abstract_kernel.F90:219-288) + kernel_device%create
 * (src/dtfft_kernel_device.F90:61-102).
 *   dims[ndims]        local extents, dims[0] fastest (ndims = 2 or 3)
 *   kernel_type        DTFFTB_KERNEL_*
 *   base_storage       bytes per element: 4, 8 or 16
 *   neighbor_data      5 x n_neighbors int32, column-major exactly like the Fortran
 *                      neighbor_data(5, P): (n1, n2, n3, in displ, out displ) per peer, in
 *                      elements; NULL for the whole-buffer permutes / KERNEL_COPY
 *   effort             dtfft_effort_t value (0..3); force_effort as in the reference
 * Zero-volume dims or KERNEL_DUMMY give a valid handle whose execute is a no-op. */
int dtfftb_kernel_create(dtfftb_kernel_t* kernel, int ndims, const int32_t* dims, int kernel_type,
                         int64_t base_storage, const int32_t* neighbor_data, int n_neighbors, int effort,
                         int force_effort);

/* abstract_kernel%execute (src/dtfft_abstract_kernel.F90:290-403) + kernel_device%execute
 * (src/dtfft_kernel_device.F90:104-178).  `stream` is a cudaStream_t.  `neighbor` is
 * 1-based and required (> 0) for the per-peer kinds (*_PIPELINED, PACK_FORWARD,
 * PACK_BACKWARD); pass 0 otherwise.  All peers of the non-pipelined kinds are covered by
 * ONE launch.  `sync` != 0 synchronises the stream after the launch. */
int dtfftb_kernel_execute(dtfftb_kernel_t kernel, const void* in, void* out, void* stream, int neighbor, int sync);

/* Extension used by the fused P2P backend: run the per-peer kind for EVERY peer in one
 * launch (what the reference does with P launches of pack_forward/pack_backward in its
 * fused backends, src/dtfft_backend_mpi.F90:474-624). */
int dtfftb_kernel_execute_all(dtfftb_kernel_t kernel, const void* in, void* out, void* stream);

/* Extension: per-peer destination bases (peer-mapped device pointers).  When set, block n
 * is written at out_bases[n] + out displacement instead of `out` (pack fused into the
 * remote NVLink store).  Pass NULL to clear.  `out_displs_override` (elements, may be
 * NULL) replaces neighbor_data(5, n) -- the receive displacement on the peer. */
int dtfftb_kernel_set_peer_out(dtfftb_kernel_t kernel, void* const* out_bases, const int64_t* out_displs_override);

/* abstract_kernel%destroy.  Sets *kernel to NULL. */
int dtfftb_kernel_destroy(dtfftb_kernel_t* kernel);

/* A kernel over explicit boxes -- what the plan layer's direct-store (NVLINK_FUSED) transpositions and brick
 * reshapes run (the reference's fused backends drive pack_forward / pack_backward per peer from
 * reshape_handle_generic, src/dtfft_reshape_handle_generic.F90:447; here one launch covers every peer).
 * `family` 2 = tiled transpose (input contiguous along a, output along b), 3 = row copy; boxes = 10 x n int64:
 * n0 n1 n2 in_off out_off is1 is2 os0 os1 os2 in elements.  `out_bases` (n pointers or NULL): box i is written
 * relative to out_bases[i] (a peer-mapped buffer) instead of the launch's `out`; NULL entries mean `out`.
 * Run it with dtfftb_kernel_execute_all, or box i alone with dtfftb_kernel_execute(neighbor = i + 1). */
int dtfftb_kernel_create_boxes(dtfftb_kernel_t* kernel, int family, int64_t base_storage, int n_boxes,
                               const int64_t* boxes, void* const* out_bases);

/* Introspection (used by tests, autotune and `report`). */
int dtfftb_kernel_get_info(dtfftb_kernel_t kernel, int* family /*0 none,1 copy,2 transpose,3 rows*/,
                           int* unit_bytes, int* tile_a, int* tile_b, int* threads, int64_t* n_items);
/* Override the family-T tile (ka, kb in multiples of 32 elements; rows = threadIdx.y extent). */
int dtfftb_kernel_set_tile(dtfftb_kernel_t kernel, int ka, int kb, int rows);
/* Time every supported tile configuration on (in, out) and keep the fastest -- the
 * reference's timed kernel autotune (src/dtfft_kernel_device.F90:338-397). Returns best ms. */
int dtfftb_kernel_autotune(dtfftb_kernel_t kernel, const void* in, void* out, void* stream, int n_warmup,
                           int n_iters, float* best_ms);
/* Same, and reports every candidate it timed like the reference's log line
 * (src/dtfft_kernel_device.F90:385-389: time and bandwidth per candidate): tiles[3*i..] = (tile_a, tile_b,
 * threads) in elements / threads, ms[i], gbs[i] = 2 x bytes moved / time.  At most max_entries are
 * written; *n_entries is the number of candidates timed. */
int dtfftb_kernel_autotune_report(dtfftb_kernel_t kernel, const void* in, void* out, void* stream, int n_warmup,
                                  int n_iters, int max_entries, int* n_entries, int32_t* tiles, float* ms, double* gbs);

/* ---- host-only introspection (tests on CPU boxes; nothing here touches a device) ----------
 * A "dry" kernel builds its geometry and its device tables exactly as a real one but never
 * uploads or launches them (execute returns DTFFT_ERROR_GPU_NOT_SET).
 * dtfftb_kernel_create_boxes_dry: kernel over explicit boxes like the plan layer's fused NVLink
 * path and brick reshapes; `family` 2 = tiled transpose, 3 = row copy; boxes = 10 x n int64
 * (n0 n1 n2 in_off out_off is1 is2 os0 os1 os2, elements); `remote_peers` != 0 gives box i its
 * own destination buffer (stand-in for a peer-mapped pointer; turns on the peer interleaving).
 * dtfftb_kernel_dump_table: the work-item table of one launch -- all peers (`neighbor` = 0) or one
 * (1-based); `unit` = 4 / 8 / 16 selects the access width of the row-copy family (ignored by the
 * transpose family).  rows = 20 x cap int64 per block: in_off out_off is1 is2 os0 os1 os2
 * item_begin shuffle n0 n1 n2 tiles0 tiles1 div0.mul div0.shr div1.mul div1.shr dest bshift
 * (offsets / strides in `unit`s for the row-copy family, in elements for the transpose family;
 * dest = index of the destination buffer or -1 for the launch's `out`; bshift = elements by which the tile grid of the
 * transpose family starts before the box along b, so that tile boundaries fall on 128-byte lines of the destination).  launch[3] = (KA, KB, ROWS)
 * of transpose_tiles_kernel or (TX, TY, rows per thread) of rows_copy_kernel. */
int dtfftb_kernel_create_dry(dtfftb_kernel_t* kernel, int ndims, const int32_t* dims, int kernel_type,
                             int64_t base_storage, const int32_t* neighbor_data, int n_neighbors);
int dtfftb_kernel_create_boxes_dry(dtfftb_kernel_t* kernel, int family, int64_t base_storage, int n_boxes,
                                   const int64_t* boxes, int remote_peers);
int dtfftb_kernel_dump_table(dtfftb_kernel_t kernel, int unit, int neighbor, int32_t cap, int64_t* rows,
                             int32_t* n_blocks, int64_t* total_items, int32_t* launch);

/* ------------------------------------------------------------------------------------
 * Exchange-backend plugin surface -- replaces backend_nccl (src/dtfft_backend_nccl.F90:38-134)
 * behind the deferred interface of abstract_backend (create_private / execute_private /
 * destroy_private, src/dtfft_abstract_backend.F90:114-139) together with the part of
 * abstract_backend%create / execute it inherits (:143-343: float-unit conversion, aux sizing,
 * self-copy on a second stream with event fork / join).  The NCCL communicator comes from
 * dtfftb_nccl_comm_create, the replacement of backend_helper%create (:395-457).
 * ---------------------------------------------------------------------------------- */
typedef struct dtfftb_backend_s* dtfftb_backend_t;
typedef struct dtfftb_executor_s* dtfftb_executor_t;

/* One NCCL communicator over the process group `comm` (NULL = 1 rank); `*nccl_comm` is an
 * ncclComm_t.  Collective. */
int dtfftb_nccl_comm_create(const dtfftb_comm_t* comm, void** nccl_comm);
int dtfftb_nccl_comm_destroy(void** nccl_comm);

/*   backend_type   DTFFT_BACKEND_NCCL (24) or DTFFT_BACKEND_NCCL_PIPELINED (27)
 *   comm_rank/size my index in, and size of, the 1-D communicator of the transposition
 *   comm_mapping   [comm_size] member -> rank in `nccl_comm` (comm_mappings, :432-441); NULL = identity
 *   *_displs/_counts  [comm_size] in ELEMENTS of `base_storage` bytes, displacements 0-based
 *                  (the reference keeps them 1-based in float units internally, :160-183) */
int dtfftb_backend_create(dtfftb_backend_t* backend, int backend_type, void* nccl_comm, int comm_rank, int comm_size,
                          const int32_t* comm_mapping, const int64_t* send_displs, const int64_t* send_counts,
                          const int64_t* recv_displs, const int64_t* recv_counts, int64_t base_storage);
/* Pipelined flavour: the per-peer unpack kernel the backend launches as blocks arrive
 * (abstract_backend%set_unpack_kernel, :345-352).  The kernel stays owned by the caller. */
int dtfftb_backend_set_unpack_kernel(dtfftb_backend_t backend, dtfftb_kernel_t unpack_kernel);
/* Workspace the pipelined flavour needs in `aux` (0 for the plain one), :196-201. */
int dtfftb_backend_get_aux_bytes(dtfftb_backend_t backend, int64_t* aux_bytes);
/* abstract_backend%execute (:223-293) + backend_nccl%execute_private (src/dtfft_backend_nccl.F90:65-134):
 * plain: grouped ncclSend / ncclRecv all-to-all(v) `in` -> `out`; pipelined: self block copied and
 * unpacked on a second stream, peers received into `aux` and unpacked into `out` one by one. */
int dtfftb_backend_execute(dtfftb_backend_t backend, void* in, void* out, void* stream, void* aux);
int dtfftb_backend_destroy(dtfftb_backend_t* backend);

/* ------------------------------------------------------------------------------------
 * FFT-executor plugin surface -- replaces cufft_executor
 * (src/interfaces/fft/cufft/dtfft_executor_cufft_m.F90:52-125) behind the deferred interface of
 * abstract_executor (src/dtfft_abstract_executor.F90:67-112), argument for argument:
 *   fft_rank 1 or 2; fft_type 0 = c2c, 1 = r2c (2 = r2r -> DTFFT_ERROR_R2R_FFT_NOT_SUPPORTED);
 *   precision dtfft_precision_t; idist / odist elements between consecutive transforms of the
 *   input / output; how_many transforms; fft_sizes / inembed / onembed [fft_rank], slowest first.
 * Batched along the fastest dimension, unit stride, unnormalised; execute(a, b, sign): sign -1
 * forward, +1 backward (c2r for r2c plans).  how_many == 0 gives a valid no-op handle.
 * ---------------------------------------------------------------------------------- */
int dtfftb_executor_create(dtfftb_executor_t* executor, int fft_rank, int fft_type, int precision, int32_t idist,
                           int32_t odist, int32_t how_many, const int32_t* fft_sizes, const int32_t* inembed,
                           const int32_t* onembed, void* stream);
int dtfftb_executor_execute(dtfftb_executor_t executor, void* a, void* b, int sign);
int dtfftb_executor_destroy(dtfftb_executor_t* executor);

/* Library info */
const char* dtfftb_version(void);
/* 1 if a CUDA device is usable in this process, else 0 (never falls back to CPU). */
int dtfftb_device_available(void);

#ifdef __cplusplus
}
#endif
#endif /* DTFFT_B200_H */
