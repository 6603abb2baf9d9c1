/*
 * dtfft_b200_api.h -- the public dtFFT plan API, served by libdtfft_b200.so.
 *
 * Function names, argument order, enum values, struct layouts and error codes are those of
 * the reference's public C header (include/dtfft.h:397-1407 and include/dtfft_config.h.in:
 * 35-193 of ShatrovOA/dtFFT v3.2.0, built WITH_CUDA), so that C / C++ / Fortran callers of
 * `dtfft_create_plan_*`, `dtfft_execute`, `dtfft_transpose`, `dtfft_reshape`,
 * `dtfft_get_local_sizes`, `dtfft_mem_alloc` ... relink unchanged.  Two deliberate
 * differences, both forced by this build having no MPI:
 *
 *   1. The communicator argument is `dtfft_comm_t` (= const dtfftb_comm_t*): a rank, a size
 *      and ONE host collective (allgather of N bytes per rank).  NULL means a single rank.
 *      With MPI present, include dtfft_b200_mpi.h, which converts an MPI_Comm.
 *      The reference uses the communicator for exactly this metadata exchange
 *      (src/dtfft_reshape_handle_generic.F90:143-144, src/dtfft_abstract_backend.F90:437-441).
 *   2. The only platform is CUDA (DTFFT_PLATFORM_CUDA is the default; HOST is rejected with
 *      DTFFT_ERROR_INVALID_PLATFORM): there is no CPU path in this library.
 *
 * Buffers passed to execute / transpose / reshape are device pointers; work is enqueued on
 * the plan stream (dtfft_get_stream) and the call returns immediately, like the reference.
 * As in the reference's GPU build, the contents of `in` are destroyed by transpose / reshape
 * (src/dtfft_plan.F90:352-353) -- except with DTFFT_BACKEND_NVLINK_FUSED, which leaves it intact.
 */
#ifndef DTFFT_B200_API_H
#define DTFFT_B200_API_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#include "dtfft_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DTFFT_VERSION_MAJOR 3
#define DTFFT_VERSION_MINOR 2
#define DTFFT_VERSION_PATCH 0
#define DTFFT_VERSION(X, Y, Z) ((X)*100000 + (Y)*1000 + (Z))
#define DTFFT_VERSION_CODE DTFFT_VERSION(DTFFT_VERSION_MAJOR, DTFFT_VERSION_MINOR, DTFFT_VERSION_PATCH)

typedef const dtfftb_comm_t* dtfft_comm_t;

/* include/dtfft_config.h.in:82-151 */
typedef enum {
    DTFFT_SUCCESS = 0,
    DTFFT_ERROR_MPI_FINALIZED = -1,
    DTFFT_ERROR_PLAN_NOT_CREATED = 1,
    DTFFT_ERROR_INVALID_TRANSPOSE_TYPE = 2,
    DTFFT_ERROR_INVALID_N_DIMENSIONS = 3,
    DTFFT_ERROR_INVALID_DIMENSION_SIZE = 4,
    DTFFT_ERROR_INVALID_COMM_TYPE = 5,
    DTFFT_ERROR_INVALID_PRECISION = 6,
    DTFFT_ERROR_INVALID_EFFORT = 7,
    DTFFT_ERROR_INVALID_EXECUTOR = 8,
    DTFFT_ERROR_INVALID_COMM_DIMS = 9,
    DTFFT_ERROR_INVALID_COMM_FAST_DIM = 10,
    DTFFT_ERROR_MISSING_R2R_KINDS = 11,
    DTFFT_ERROR_INVALID_R2R_KINDS = 12,
    DTFFT_ERROR_R2C_TRANSPOSE_PLAN = 13,
    DTFFT_ERROR_INPLACE_TRANSPOSE = 14,
    DTFFT_ERROR_INVALID_AUX = 15,
    DTFFT_ERROR_INVALID_LAYOUT = 16,
    DTFFT_ERROR_INVALID_USAGE = 17,
    DTFFT_ERROR_PLAN_IS_CREATED = 18,
    DTFFT_ERROR_ALLOC_FAILED = 19,
    DTFFT_ERROR_FREE_FAILED = 20,
    DTFFT_ERROR_INVALID_ALLOC_BYTES = 21,
    DTFFT_ERROR_DLOPEN_FAILED = 22,
    DTFFT_ERROR_DLSYM_FAILED = 23,
    DTFFT_ERROR_PENCIL_ARRAYS_SIZE_MISMATCH = 25,
    DTFFT_ERROR_PENCIL_ARRAYS_INVALID_SIZES = 26,
    DTFFT_ERROR_PENCIL_INVALID_COUNTS = 27,
    DTFFT_ERROR_PENCIL_INVALID_STARTS = 28,
    DTFFT_ERROR_PENCIL_SHAPE_MISMATCH = 29,
    DTFFT_ERROR_PENCIL_OVERLAP = 30,
    DTFFT_ERROR_PENCIL_NOT_CONTINUOUS = 31,
    DTFFT_ERROR_PENCIL_NOT_INITIALIZED = 32,
    DTFFT_ERROR_INVALID_MEASURE_WARMUP_ITERS = 33,
    DTFFT_ERROR_INVALID_MEASURE_ITERS = 34,
    DTFFT_ERROR_INVALID_REQUEST = 35,
    DTFFT_ERROR_TRANSPOSE_ACTIVE = 36,
    DTFFT_ERROR_TRANSPOSE_NOT_ACTIVE = 37,
    DTFFT_ERROR_INVALID_RESHAPE_TYPE = 38,
    DTFFT_ERROR_RESHAPE_ACTIVE = 39,
    DTFFT_ERROR_RESHAPE_NOT_ACTIVE = 40,
    DTFFT_ERROR_INPLACE_RESHAPE = 41,
    DTFFT_ERROR_INVALID_EXECUTE_TYPE = 43,
    DTFFT_ERROR_RESHAPE_NOT_SUPPORTED = 44,
    DTFFT_ERROR_R2C_EXECUTE_CALLED = 45,
    DTFFT_ERROR_INVALID_CART_COMM = 46,
    DTFFT_ERROR_INVALID_TRANSPOSE_MODE = 47,
    DTFFT_ERROR_INVALID_ACCESS_MODE = 48,
    DTFFT_ERROR_R2R_FFT_NOT_SUPPORTED = 101,
    DTFFT_ERROR_GPU_INVALID_STREAM = 201,
    DTFFT_ERROR_INVALID_BACKEND = 202,
    DTFFT_ERROR_GPU_NOT_SET = 203,
    DTFFT_ERROR_VKFFT_R2R_2D_PLAN = 204,
    DTFFT_ERROR_BACKENDS_DISABLED = 205,
    DTFFT_ERROR_NOT_DEVICE_PTR = 300,
    DTFFT_ERROR_NOT_NVSHMEM_PTR = 301,
    DTFFT_ERROR_INVALID_PLATFORM = 400,
    DTFFT_ERROR_INVALID_PLATFORM_EXECUTOR = 401,
    DTFFT_ERROR_INVALID_PLATFORM_BACKEND = 402
} dtfft_error_t;

typedef enum { DTFFT_EXECUTE_FORWARD = 11, DTFFT_EXECUTE_BACKWARD = 12 } dtfft_execute_t;

typedef enum {
    DTFFT_TRANSPOSE_X_TO_Y = 1,
    DTFFT_TRANSPOSE_Y_TO_X = -1,
    DTFFT_TRANSPOSE_Y_TO_Z = 2,
    DTFFT_TRANSPOSE_Z_TO_Y = -2,
    DTFFT_TRANSPOSE_X_TO_Z = 3,
    DTFFT_TRANSPOSE_Z_TO_X = -3
} dtfft_transpose_t;

typedef enum {
    DTFFT_RESHAPE_X_BRICKS_TO_PENCILS = 11,
    DTFFT_RESHAPE_X_PENCILS_TO_BRICKS = 12,
    DTFFT_RESHAPE_Z_PENCILS_TO_BRICKS = 13,
    DTFFT_RESHAPE_Z_BRICKS_TO_PENCILS = 14,
    DTFFT_RESHAPE_Y_BRICKS_TO_PENCILS = 14,
    DTFFT_RESHAPE_Y_PENCILS_TO_BRICKS = 13
} dtfft_reshape_t;

typedef enum { DTFFT_SINGLE = 0, DTFFT_DOUBLE = 1 } dtfft_precision_t;
typedef enum { DTFFT_ESTIMATE = 0, DTFFT_MEASURE = 1, DTFFT_PATIENT = 2, DTFFT_EXHAUSTIVE = 3 } dtfft_effort_t;
typedef enum {
    DTFFT_EXECUTOR_NONE = 0,
    DTFFT_EXECUTOR_FFTW3 = 1,
    DTFFT_EXECUTOR_MKL = 2,
    DTFFT_EXECUTOR_CUFFT = 3,
    DTFFT_EXECUTOR_VKFFT = 4
} dtfft_executor_t;
typedef enum {
    DTFFT_DCT_1 = 3,
    DTFFT_DCT_2 = 5,
    DTFFT_DCT_3 = 4,
    DTFFT_DCT_4 = 6,
    DTFFT_DST_1 = 7,
    DTFFT_DST_2 = 9,
    DTFFT_DST_3 = 8,
    DTFFT_DST_4 = 10
} dtfft_r2r_kind_t;
typedef enum {
    DTFFT_LAYOUT_X_BRICKS = 1,
    DTFFT_LAYOUT_X_PENCILS = 2,
    DTFFT_LAYOUT_X_PENCILS_FOURIER = 3,
    DTFFT_LAYOUT_Y_PENCILS = 4,
    DTFFT_LAYOUT_Z_PENCILS = 5,
    DTFFT_LAYOUT_Z_BRICKS = 6
} dtfft_layout_t;
typedef enum {
    DTFFT_BACKEND_MPI_DATATYPE = 21,
    DTFFT_BACKEND_MPI_P2P = 22,
    DTFFT_BACKEND_MPI_A2A = 23,
    DTFFT_BACKEND_NCCL = 24,
    DTFFT_BACKEND_CUFFTMP = 25,
    DTFFT_BACKEND_MPI_P2P_PIPELINED = 26,
    DTFFT_BACKEND_NCCL_PIPELINED = 27,
    DTFFT_BACKEND_CUFFTMP_PIPELINED = 28,
    DTFFT_BACKEND_MPI_RMA = 29,
    DTFFT_BACKEND_MPI_RMA_PIPELINED = 30,
    DTFFT_BACKEND_MPI_P2P_SCHEDULED = 31,
    DTFFT_BACKEND_MPI_P2P_FUSED = 32,
    DTFFT_BACKEND_MPI_RMA_FUSED = 33,
    DTFFT_BACKEND_MPI_P2P_COMPRESSED = 34,
    DTFFT_BACKEND_MPI_RMA_COMPRESSED = 35,
    DTFFT_BACKEND_ADAPTIVE = 36,
    DTFFT_BACKEND_NCCL_COMPRESSED = 37,
    /* new in this library: one fused pack + NVLink peer store kernel per transposition */
    DTFFT_BACKEND_NVLINK_FUSED = 38,
    DTFFT_BACKEND_NONE = -111
} dtfft_backend_t;
typedef enum { DTFFT_TRANSPOSE_MODE_PACK = 15, DTFFT_TRANSPOSE_MODE_UNPACK = 16 } dtfft_transpose_mode_t;
typedef enum { DTFFT_ACCESS_MODE_WRITE = -1, DTFFT_ACCESS_MODE_READ = 1 } dtfft_access_mode_t;
typedef enum { DTFFT_PLATFORM_HOST = 1, DTFFT_PLATFORM_CUDA = 2 } dtfft_platform_t;

typedef void* dtfft_plan_t;
typedef void* dtfft_request_t;
typedef void* dtfft_stream_t; /* cudaStream_t */

/* include/dtfft.h:364-380 */
typedef struct {
    uint8_t dim;
    uint8_t ndims;
    int32_t starts[3];
    int32_t counts[3];
    size_t size;
} dtfft_pencil_t;

/* include/dtfft.h:1159-1389 (WITH_CUDA, without compression); defaults src/dtfft_config.F90:644-669 */
typedef struct {
    bool enable_log;
    bool enable_z_slab;
    bool enable_y_slab;
    int32_t n_measure_warmup_iters;
    int32_t n_measure_iters;
    dtfft_platform_t platform;
    dtfft_stream_t stream;
    dtfft_backend_t backend;
    dtfft_backend_t reshape_backend;
    bool enable_datatype_backend;
    bool enable_mpi_backends;
    bool enable_pipelined_backends;
    bool enable_rma_backends;
    bool enable_fused_backends;
    bool enable_nccl_backends;
    bool enable_nvshmem_backends;
    bool enable_kernel_autotune;
    bool enable_fourier_reshape;
    dtfft_transpose_mode_t transpose_mode;
    dtfft_access_mode_t access_mode;
} dtfft_config_t;

/* replaces dtfft_get_version, include/dtfft.h:52 */
int32_t dtfft_get_version(void);

/* constructors: include/dtfft.h:397-515 */
/* replaces dtfft_create_plan_r2r, include/dtfft.h:398 */
dtfft_error_t dtfft_create_plan_r2r(int8_t ndims, const int32_t* dims, const dtfft_r2r_kind_t* kinds, dtfft_comm_t comm,
                                    dtfft_precision_t precision, dtfft_effort_t effort, dtfft_executor_t executor,
                                    dtfft_plan_t* plan);
/* replaces dtfft_create_plan_r2r_pencil, include/dtfft.h:423 */
dtfft_error_t dtfft_create_plan_r2r_pencil(const dtfft_pencil_t* pencil, const dtfft_r2r_kind_t* kinds,
                                           dtfft_comm_t comm, dtfft_precision_t precision, dtfft_effort_t effort,
                                           dtfft_executor_t executor, dtfft_plan_t* plan);
/* replaces dtfft_create_plan_c2c, include/dtfft.h:445 */
dtfft_error_t dtfft_create_plan_c2c(int8_t ndims, const int32_t* dims, dtfft_comm_t comm, dtfft_precision_t precision,
                                    dtfft_effort_t effort, dtfft_executor_t executor, dtfft_plan_t* plan);
/* replaces dtfft_create_plan_c2c_pencil, include/dtfft.h:466 */
dtfft_error_t dtfft_create_plan_c2c_pencil(const dtfft_pencil_t* pencil, dtfft_comm_t comm,
                                           dtfft_precision_t precision, dtfft_effort_t effort,
                                           dtfft_executor_t executor, dtfft_plan_t* plan);
/* replaces dtfft_create_plan_r2c, include/dtfft.h:487 */
dtfft_error_t dtfft_create_plan_r2c(int8_t ndims, const int32_t* dims, dtfft_comm_t comm, dtfft_precision_t precision,
                                    dtfft_effort_t effort, dtfft_executor_t executor, dtfft_plan_t* plan);
/* replaces dtfft_create_plan_r2c_pencil, include/dtfft.h:509 */
dtfft_error_t dtfft_create_plan_r2c_pencil(const dtfft_pencil_t* pencil, dtfft_comm_t comm,
                                           dtfft_precision_t precision, dtfft_effort_t effort,
                                           dtfft_executor_t executor, dtfft_plan_t* plan);

/* execution: include/dtfft.h:555-645 */
/* replaces dtfft_execute, include/dtfft.h:555 */
dtfft_error_t dtfft_execute(dtfft_plan_t plan, void* in, void* out, dtfft_execute_t execute_type, void* aux);
/* replaces dtfft_transpose, include/dtfft.h:570 */
dtfft_error_t dtfft_transpose(dtfft_plan_t plan, void* in, void* out, dtfft_transpose_t transpose_type, void* aux);
/* replaces dtfft_transpose_start, include/dtfft.h:591 */
dtfft_error_t dtfft_transpose_start(dtfft_plan_t plan, void* in, void* out, dtfft_transpose_t transpose_type,
                                    void* aux, dtfft_request_t* request);
/* replaces dtfft_transpose_end, include/dtfft.h:602 */
dtfft_error_t dtfft_transpose_end(dtfft_plan_t plan, dtfft_request_t request);
/* replaces dtfft_reshape, include/dtfft.h:617 */
dtfft_error_t dtfft_reshape(dtfft_plan_t plan, void* in, void* out, dtfft_reshape_t reshape_type, void* aux);
/* replaces dtfft_reshape_start, include/dtfft.h:635 */
dtfft_error_t dtfft_reshape_start(dtfft_plan_t plan, void* in, void* out, dtfft_reshape_t reshape_type, void* aux,
                                  dtfft_request_t* request);
/* replaces dtfft_reshape_end, include/dtfft.h:645 */
dtfft_error_t dtfft_reshape_end(dtfft_plan_t plan, dtfft_request_t request);
/* replaces dtfft_destroy, include/dtfft.h:654 */
dtfft_error_t dtfft_destroy(dtfft_plan_t* plan);

/* sizes and metadata: include/dtfft.h:668-918 */
/* replaces dtfft_get_local_sizes, include/dtfft.h:668 */
dtfft_error_t dtfft_get_local_sizes(dtfft_plan_t plan, int32_t* in_starts, int32_t* in_counts, int32_t* out_starts,
                                    int32_t* out_counts, size_t* alloc_size);
/* replaces dtfft_get_alloc_size, include/dtfft.h:678 */
dtfft_error_t dtfft_get_alloc_size(dtfft_plan_t plan, size_t* alloc_size);
/* replaces dtfft_get_aux_size, include/dtfft.h:688 */
dtfft_error_t dtfft_get_aux_size(dtfft_plan_t plan, size_t* aux_size);
/* replaces dtfft_get_aux_bytes, include/dtfft.h:698 */
dtfft_error_t dtfft_get_aux_bytes(dtfft_plan_t plan, size_t* aux_bytes);
/* replaces dtfft_get_aux_size_reshape, include/dtfft.h:708 */
dtfft_error_t dtfft_get_aux_size_reshape(dtfft_plan_t plan, size_t* aux_size);
/* replaces dtfft_get_aux_bytes_reshape, include/dtfft.h:718 */
dtfft_error_t dtfft_get_aux_bytes_reshape(dtfft_plan_t plan, size_t* aux_bytes);
/* replaces dtfft_get_aux_size_transpose, include/dtfft.h:728 */
dtfft_error_t dtfft_get_aux_size_transpose(dtfft_plan_t plan, size_t* aux_size);
/* replaces dtfft_get_aux_bytes_transpose, include/dtfft.h:738 */
dtfft_error_t dtfft_get_aux_bytes_transpose(dtfft_plan_t plan, size_t* aux_bytes);
/* replaces dtfft_get_pencil, include/dtfft.h:806 */
dtfft_error_t dtfft_get_pencil(dtfft_plan_t plan, dtfft_layout_t layout, dtfft_pencil_t* pencil);
/* replaces dtfft_get_element_size, include/dtfft.h:817 */
dtfft_error_t dtfft_get_element_size(dtfft_plan_t plan, size_t* element_size);
/* replaces dtfft_get_alloc_bytes, include/dtfft.h:831 */
dtfft_error_t dtfft_get_alloc_bytes(dtfft_plan_t plan, size_t* alloc_bytes);
/* replaces dtfft_mem_alloc, include/dtfft.h:843 */
dtfft_error_t dtfft_mem_alloc(dtfft_plan_t plan, size_t alloc_bytes, void** ptr);
/* replaces dtfft_mem_free, include/dtfft.h:854 */
dtfft_error_t dtfft_mem_free(dtfft_plan_t plan, void* ptr);
/* replaces dtfft_report, include/dtfft.h:864 */
dtfft_error_t dtfft_report(dtfft_plan_t plan);
/* replaces dtfft_get_z_slab_enabled, include/dtfft.h:526 */
dtfft_error_t dtfft_get_z_slab_enabled(dtfft_plan_t plan, bool* is_z_slab_enabled);
/* replaces dtfft_get_y_slab_enabled, include/dtfft.h:537 */
dtfft_error_t dtfft_get_y_slab_enabled(dtfft_plan_t plan, bool* is_y_slab_enabled);
/* replaces dtfft_get_executor, include/dtfft.h:875 */
dtfft_error_t dtfft_get_executor(dtfft_plan_t plan, dtfft_executor_t* executor);
/* replaces dtfft_get_precision, include/dtfft.h:886 */
dtfft_error_t dtfft_get_precision(dtfft_plan_t plan, dtfft_precision_t* precision);
/* replaces dtfft_get_dims, include/dtfft.h:901 */
dtfft_error_t dtfft_get_dims(dtfft_plan_t plan, int8_t* ndims, const int32_t* dims[]);
/* replaces dtfft_get_grid_dims, include/dtfft.h:917 */
dtfft_error_t dtfft_get_grid_dims(dtfft_plan_t plan, int8_t* ndims, const int32_t* grid_dims[]);
/* replaces dtfft_get_stream, include/dtfft.h:1096 */
dtfft_error_t dtfft_get_stream(dtfft_plan_t plan, dtfft_stream_t* stream);
/* replaces dtfft_get_platform, include/dtfft.h:1107 */
dtfft_error_t dtfft_get_platform(dtfft_plan_t plan, dtfft_platform_t* platform);
/* replaces dtfft_get_backend, include/dtfft.h:1122 */
dtfft_error_t dtfft_get_backend(dtfft_plan_t plan, dtfft_backend_t* backend);
/* replaces dtfft_get_reshape_backend, include/dtfft.h:1133 */
dtfft_error_t dtfft_get_reshape_backend(dtfft_plan_t plan, dtfft_backend_t* backend);
/* replaces dtfft_get_backend_pipelined, include/dtfft.h:1154 */
dtfft_error_t dtfft_get_backend_pipelined(const dtfft_backend_t backend, bool* is_pipe);

/* replaces dtfft_get_error_string, include/dtfft.h:759 */
const char* dtfft_get_error_string(dtfft_error_t error_code);
/* replaces dtfft_get_precision_string, include/dtfft.h:768 */
const char* dtfft_get_precision_string(dtfft_precision_t precision);
/* replaces dtfft_get_executor_string, include/dtfft.h:777 */
const char* dtfft_get_executor_string(dtfft_executor_t executor);
/* replaces dtfft_get_backend_string, include/dtfft.h:1143 */
const char* dtfft_get_backend_string(dtfft_backend_t backend);

/* configuration: include/dtfft.h:1398-1407 */
/* replaces dtfft_create_config, include/dtfft.h:1398 */
dtfft_error_t dtfft_create_config(dtfft_config_t* config);
/* replaces dtfft_set_config, include/dtfft.h:1407 */
dtfft_error_t dtfft_set_config(const dtfft_config_t* config);

/* ---- extensions of this library (not in the reference) --------------------------------- */
/* DTFFT_BACKEND_NVLINK_FUSED takes ANY device pointer (like the reference, src/dtfft_plan.F90:1769-1795): the first call
 * with a given destination publishes the allocation behind it to the peers (collective, over cudaIpc), a re-allocated
 * address is detected and re-published, memory cudaIpc cannot share runs on an NCCL stand-in.  These two calls are
 * therefore OPTIONAL: register_buffer maps a buffer ahead of its first use (collective: every rank calls it in the same
 * order with its own buffer of the same role; dtfft_mem_alloc does it automatically); unregister_buffer (collective) makes
 * the peers drop their mappings before the memory is freed -- what dtfft_mem_free does for dtfft_mem_alloc'ed memory, and
 * what a caller should do (or let the buffer outlive the plan) for memory it frees itself. */
dtfft_error_t dtfftb_plan_register_buffer(dtfft_plan_t plan, void* ptr, size_t bytes);
dtfft_error_t dtfftb_plan_unregister_buffer(dtfft_plan_t plan, void* ptr);
/* Per-execute accounting of the last dtfft_execute / dtfft_transpose / dtfft_reshape on this
 * rank: kernels of this library launched, payload bytes moved by them (one direction) and
 * bytes that left the GPU. */
dtfft_error_t dtfftb_plan_get_stats(dtfft_plan_t plan, int64_t* kernel_launches, int64_t* local_bytes,
                                    int64_t* remote_bytes);
/* 0 if healthy; non-zero if a device barrier of the NVLink backend timed out. */
int dtfftb_plan_peer_error(dtfft_plan_t plan);
/* Stage overlap of the cuFFT executor with the NVLINK_FUSED exchange inside dtfft_execute
 * (extension; the reference serialises FFT and exchange on one stream, src/dtfft_plan.F90:1057-1101):
 * the FFT before a transposition is cut into `nchunks` ranges of its slowest axis and chunk k is
 * stored to the peers on a second stream while chunk k+1 is transformed.  nchunks <= 1 disables;
 * `exchange_ctas` = CTAs of the persistent exchange kernel (0 = one per SM).  Default: a plan-wide
 * rule at DTFFT_ESTIMATE (8 chunks when the longest transform has >= 4096 points and the local array
 * is >= 256 MiB, else off), a timed choice among {1, 4, 8} at effort >= DTFFT_MEASURE; the env
 * variables DTFFTB_OVERLAP_CHUNKS / DTFFTB_OVERLAP_CTAS and this call override both.  Must be set
 * identically on every rank. */
dtfft_error_t dtfftb_plan_set_overlap(dtfft_plan_t plan, int nchunks, int exchange_ctas);
dtfft_error_t dtfftb_plan_get_overlap(dtfft_plan_t plan, int* nchunks);
/* dtfft_execute replays a CUDA graph: the first call with a given (in, out, aux, direction) runs
 * eagerly, the second is captured while it is enqueued, later ones are one cudaGraphLaunch
 * (extension; single-GPU and NVLINK_FUSED plans, env DTFFTB_GRAPHS=0 disables).  Freeing or
 * unregistering a buffer drops the graphs. */
dtfft_error_t dtfftb_plan_set_graphs(dtfft_plan_t plan, int enable);
dtfft_error_t dtfftb_plan_get_graph_replays(dtfft_plan_t plan, int64_t* n_replays);
/* How one transposition of the plan moves its data on this rank: *form = 0 local kernel only (one rank in its
 * communicator), 1 NCCL (pack -> ncclSend/Recv -> unpack), 2 NVLINK_FUSED direct-store kernel, 3 NVLINK_FUSED copy-engine
 * form (pack + one strided 3-D copy per peer slice; *n_slices = copies per execute), 4 direct-store kernel when run on
 * its own and the copy-engine form when pipelined with the local transposition next to it in dtfft_execute. */
dtfft_error_t dtfftb_plan_get_exchange_form(dtfft_plan_t plan, int transpose_type, int* form, int* n_slices);
/* NVLINK_FUSED: how many transpositions / reshapes of this plan ran on the NCCL stand-in because a caller's
 * buffer could not be shared through cudaIpc (stream-ordered or virtual-memory allocations). */
dtfft_error_t dtfftb_plan_get_fallbacks(dtfft_plan_t plan, int64_t* n_fallbacks);
/* Number of FFT+transposition stages of the last dtfft_execute that ran overlapped. */
dtfft_error_t dtfftb_plan_get_overlapped_stages(dtfft_plan_t plan, int64_t* n_stages);

/* Host-metadata-only plan: decomposition, pencils, sizes and exchange geometry are computed
 * exactly as for a real plan, but no device is touched; execute / transpose / reshape /
 * mem_alloc return DTFFT_ERROR_GPU_NOT_SET.  `kind` 0 = c2c, 1 = r2c, 2 = r2r; pass `dims`
 * (with ndims) or `pencil`.  Used to test the host logic on CPU-only boxes. */
dtfft_error_t dtfftb_plan_create_dry(int kind, int8_t ndims, const int32_t* dims, const dtfft_pencil_t* pencil,
                                     dtfft_comm_t comm, dtfft_precision_t precision, dtfft_executor_t executor,
                                     dtfft_plan_t* plan);
/* Process grids 1 x g1 x g2 the DTFFT_MEASURE / DTFFT_PATIENT grid search times for a default 3-D
 * decomposition (autotune_grid_decomposition + the validity rule of autotune_grid,
 * src/dtfft_transpose_plan.F90:456-500, 600-607), in search order: pairs (g1, g2) written to
 * `grids` (up to `cap` pairs); returns their number.  Host-only. */
int32_t dtfftb_grid_candidates(const int32_t* dims, int32_t comm_size, int32_t cap, int32_t* grids);
/* Dry plans only: re-decompose a default 3-D plan on the grid 1 x g1 x g2 exactly as the grid
 * search does between two timings (test hook for the host logic). */
dtfft_error_t dtfftb_plan_dry_set_grid(dtfft_plan_t plan, int32_t g1, int32_t g2);
/* Exchange geometry of transposition / reshape `type` on this rank (works on dry and real plans).
 * All output arrays are optional (NULL) and hold `cap` peers at most; *n_members is always set.
 *   members[P]        world ranks of the 1-D communicator
 *   kernels[2]        pack / unpack kernel_type_t (transposes; reference geometry)
 *   send_nd, recv_nd  5 x P neighbor_data (src/dtfft_reshape_handle_generic.F90:406-413, 575-612)
 *   counts_displs     4 x P int64: send counts, send displs, recv counts, recv displs (elements)
 *   fused_boxes       10 x P int64 per peer: n0 n1 n2 in_off out_off is1 is2 os0 os1 os2 -- the
 *                     part of my source array stored straight into peer p's destination array
 *   fused_transposing 1 if the fused boxes change the fastest axis (family T) */
dtfft_error_t dtfftb_plan_describe_exchange(dtfft_plan_t plan, int type, int32_t cap, int32_t* n_members,
                                            int32_t* my_index, int32_t* members, int32_t* kernels, int32_t* send_nd,
                                            int32_t* recv_nd, int64_t* counts_displs, int64_t* fused_boxes,
                                            int32_t* fused_transposing);
/* Brick <-> pencil reshape `reshape_type` over the NCCL backends on this rank (dry and real plans):
 * the pack -> all-to-all(v) -> unpack geometry of reshape_handle_generic for reshapes
 * (src/dtfft_reshape_handle_generic.F90:343-376, 536-567) derived from global-index intersections.
 *   pack_boxes    10 x P int64 (layout as fused_boxes): my source -> slot p at send displ p
 *   unpack_boxes  10 x P int64: slot p at recv displ p -> my destination
 *   counts_displs 4 x P int64: send counts, send displs, recv counts, recv displs (elements)
 *   flags[3]      is_pack_free, is_unpack_free (:261-266, 479-484), reshape_strat (:267-289) */
dtfft_error_t dtfftb_plan_describe_reshape(dtfft_plan_t plan, int reshape_type, int32_t cap, int32_t* n_members,
                                           int32_t* my_index, int32_t* members, int64_t* pack_boxes,
                                           int64_t* unpack_boxes, int64_t* counts_displs, int32_t* flags);
/* Stage overlap of the cuFFT executor with the NVLINK_FUSED exchange: chunk `k` of `nchunks` (cut along the
 * slowest axis of the source pencil) of transposition `transpose_type` as a transposition of its own.
 * boxes = 10 x P int64 (layout as fused_boxes, in_off relative to the chunk), *chunk_offset = first element
 * of the chunk in the source pencil.  Works on dry and real plans. */
dtfft_error_t dtfftb_plan_describe_chunk(dtfft_plan_t plan, int transpose_type, int32_t k, int32_t nchunks, int32_t cap,
                                         int32_t* n_members, int64_t* boxes, int64_t* chunk_offset);
/* Copy-engine form of one transposition on this rank (NVLINK_FUSED, DMA mode): one entry of 30 int64 per (member,
 * slice) -- pack box (n0 n1 n2 in_off out_off is1 is2 os0 os1 os2: my source -> staging, the slice packed in the order
 * of its destination rows), then run rows planes dst_off dst_pitch dst_plane_rows ok (one strided 3-D copy: `planes` x
 * `rows` rows of `run` elements from the dense staging block to dst_off + row * dst_pitch + plane * dst_plane_rows *
 * dst_pitch of the member's destination), then the direct-store box of the same slice (10 values), then member index,
 * slice, slices of the block.  Host tests replay it. */
dtfft_error_t dtfftb_plan_describe_dma(dtfft_plan_t plan, int ttype, int32_t cap_members, int32_t cap_entries,
                                       int32_t* n_members, int32_t* me, int32_t* members, int32_t* n_entries, int64_t* rows);
/* The piece of the LOCAL transposition t_local that writes (side 0) / reads (side 1) exactly slice `sub` of what travels
 * between this rank and member `peer` of the exchanging transposition t_exchange next to it: one box, 10 values as
 * above; *nsub = slices of that block. */
dtfft_error_t dtfftb_plan_describe_peer_piece(dtfft_plan_t plan, int t_local, int t_exchange, int side, int32_t peer,
                                              int32_t sub, int32_t* nsub, int64_t* box);

#ifdef __cplusplus
}
#endif
#endif /* DTFFT_B200_API_H */
