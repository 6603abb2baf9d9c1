// Public C ABI of the plan layer (include/dtfft_b200_api.h): the symbols the reference's
// C / C++ / Fortran bindings call (src/interfaces/api/dtfft_api_c.c:26-461 -> dtfft_api.F90).
// NULL handling follows dtfft_api_c.c: a NULL plan or NULL out-pointer is
// DTFFT_ERROR_INVALID_USAGE / DTFFT_ERROR_PLAN_NOT_CREATED.
#include <mutex>
#include <new>
#include <unordered_set>

#include "errors.h"
#include "plan.h"

using dtfftb::Plan;

namespace {
inline Plan* P(dtfft_plan_t p) { return static_cast<Plan*>(p); }
inline dtfft_error_t E(int rc) { return static_cast<dtfft_error_t>(rc); }

dtfft_error_t create_any(dtfftb::PlanKind kind, int ndims, const int32_t* dims, const dtfft_pencil_t* pencil,
                         const dtfft_r2r_kind_t* kinds, dtfft_comm_t comm, dtfft_precision_t precision,
                         dtfft_effort_t effort, dtfft_executor_t executor, dtfft_plan_t* plan) {
    if (!plan) return DTFFT_ERROR_INVALID_USAGE;
    *plan = nullptr;
    if (!dims && !pencil) return DTFFT_ERROR_INVALID_USAGE;
    Plan* p = new (std::nothrow) Plan;
    if (!p) return DTFFT_ERROR_ALLOC_FAILED;
    int k[3] = {-1, -1, -1};
    if (kinds) {
        const int n = dims ? ndims : (pencil ? pencil->ndims : 0);
        for (int i = 0; i < n && i < 3; ++i) k[i] = (int)kinds[i];
    }
    int rc = p->create(kind, ndims, dims, pencil, kinds ? k : nullptr, comm, (int)precision, (int)effort, (int)executor);
    if (rc) {
        delete p;
        return E(rc);
    }
    *plan = p;
    return DTFFT_SUCCESS;
}
}  // namespace

#define PLAN_OR_RETURN(p)                                  \
    if (!(p)) return DTFFT_ERROR_PLAN_NOT_CREATED;         \
    if (!P(p)->created()) return DTFFT_ERROR_PLAN_NOT_CREATED;

extern "C" {

int32_t dtfft_get_version(void) { return DTFFT_VERSION_CODE; }

dtfft_error_t dtfft_create_plan_r2r(int8_t ndims, const int32_t* dims, const dtfft_r2r_kind_t* kinds, dtfft_comm_t comm,
                                    dtfft_precision_t precision, dtfft_effort_t effort, dtfft_executor_t executor,
                                    dtfft_plan_t* plan) {
    if (!dims) return DTFFT_ERROR_INVALID_USAGE;
    return create_any(dtfftb::PLAN_R2R, ndims, dims, nullptr, kinds, comm, precision, effort, executor, plan);
}
dtfft_error_t dtfft_create_plan_r2r_pencil(const dtfft_pencil_t* pencil, const dtfft_r2r_kind_t* kinds,
                                           dtfft_comm_t comm, dtfft_precision_t precision, dtfft_effort_t effort,
                                           dtfft_executor_t executor, dtfft_plan_t* plan) {
    if (!pencil) return DTFFT_ERROR_INVALID_USAGE;
    return create_any(dtfftb::PLAN_R2R, 0, nullptr, pencil, kinds, comm, precision, effort, executor, plan);
}
dtfft_error_t dtfft_create_plan_c2c(int8_t ndims, const int32_t* dims, dtfft_comm_t comm, dtfft_precision_t precision,
                                    dtfft_effort_t effort, dtfft_executor_t executor, dtfft_plan_t* plan) {
    if (!dims) return DTFFT_ERROR_INVALID_USAGE;
    return create_any(dtfftb::PLAN_C2C, ndims, dims, nullptr, nullptr, comm, precision, effort, executor, plan);
}
dtfft_error_t dtfft_create_plan_c2c_pencil(const dtfft_pencil_t* pencil, dtfft_comm_t comm,
                                           dtfft_precision_t precision, dtfft_effort_t effort,
                                           dtfft_executor_t executor, dtfft_plan_t* plan) {
    if (!pencil) return DTFFT_ERROR_INVALID_USAGE;
    return create_any(dtfftb::PLAN_C2C, 0, nullptr, pencil, nullptr, comm, precision, effort, executor, plan);
}
dtfft_error_t dtfft_create_plan_r2c(int8_t ndims, const int32_t* dims, dtfft_comm_t comm, dtfft_precision_t precision,
                                    dtfft_effort_t effort, dtfft_executor_t executor, dtfft_plan_t* plan) {
    if (!dims) return DTFFT_ERROR_INVALID_USAGE;
    return create_any(dtfftb::PLAN_R2C, ndims, dims, nullptr, nullptr, comm, precision, effort, executor, plan);
}
dtfft_error_t dtfft_create_plan_r2c_pencil(const dtfft_pencil_t* pencil, dtfft_comm_t comm,
                                           dtfft_precision_t precision, dtfft_effort_t effort,
                                           dtfft_executor_t executor, dtfft_plan_t* plan) {
    if (!pencil) return DTFFT_ERROR_INVALID_USAGE;
    return create_any(dtfftb::PLAN_R2C, 0, nullptr, pencil, nullptr, comm, precision, effort, executor, plan);
}

dtfft_error_t dtfft_execute(dtfft_plan_t plan, void* in, void* out, dtfft_execute_t execute_type, void* aux) {
    PLAN_OR_RETURN(plan);
    if (!in || !out) return DTFFT_ERROR_INVALID_USAGE;
    return E(P(plan)->execute(in, out, (int)execute_type, aux));
}
dtfft_error_t dtfft_transpose(dtfft_plan_t plan, void* in, void* out, dtfft_transpose_t transpose_type, void* aux) {
    PLAN_OR_RETURN(plan);
    if (!in || !out) return DTFFT_ERROR_INVALID_USAGE;
    return E(P(plan)->transpose(in, out, (int)transpose_type, aux));
}
// On the GPU every backend enqueues the whole transposition on the plan stream, so *_start
// does the work and *_end only validates and retires the request (reshape_handle_generic.F90:661-664:
// async is supported for host MPI backends only; get_async_active is .false. for the NCCL backends,
// abstract_backend.F90:345-349, so the *_ACTIVE errors cannot occur here).  A request is the
// reference's async_request (dtfft_plan.F90:62-72): operation type + buffers, checked by
// CHECK_REQUEST (:75-84) -- a null, foreign, already retired or wrong-kind request is
// DTFFT_ERROR_INVALID_REQUEST.  Live requests are kept in a registry so that a stale handle is
// rejected instead of dereferenced.
namespace {
struct AsyncRequest {
    dtfft_plan_t plan;
    int type;
    void *in, *out;
};
std::mutex g_req_mutex;
std::unordered_set<AsyncRequest*> g_requests;

dtfft_request_t new_request(dtfft_plan_t plan, int type, void* in, void* out) {
    AsyncRequest* r = new (std::nothrow) AsyncRequest{plan, type, in, out};
    if (!r) return nullptr;
    std::lock_guard<std::mutex> lk(g_req_mutex);
    g_requests.insert(r);
    return r;
}

dtfft_error_t end_request(dtfft_plan_t plan, dtfft_request_t request, bool transpose) {
    if (!request) return DTFFT_ERROR_INVALID_REQUEST;
    AsyncRequest* r = static_cast<AsyncRequest*>(request);
    std::lock_guard<std::mutex> lk(g_req_mutex);
    auto it = g_requests.find(r);
    if (it == g_requests.end()) return DTFFT_ERROR_INVALID_REQUEST;
    const bool is_transpose = r->type >= -3 && r->type <= 3 && r->type != 0;
    if (r->plan != plan || is_transpose != transpose || !r->in || !r->out) return DTFFT_ERROR_INVALID_REQUEST;
    g_requests.erase(it);
    delete r;
    return DTFFT_SUCCESS;
}

void drop_requests_of(dtfft_plan_t plan) {
    std::lock_guard<std::mutex> lk(g_req_mutex);
    for (auto it = g_requests.begin(); it != g_requests.end();) {
        if ((*it)->plan == plan) {
            delete *it;
            it = g_requests.erase(it);
        } else {
            ++it;
        }
    }
}
}  // namespace

dtfft_error_t dtfft_transpose_start(dtfft_plan_t plan, void* in, void* out, dtfft_transpose_t transpose_type,
                                    void* aux, dtfft_request_t* request) {
    if (!request) return DTFFT_ERROR_INVALID_USAGE;
    *request = nullptr;
    dtfft_error_t rc = dtfft_transpose(plan, in, out, transpose_type, aux);
    if (rc != DTFFT_SUCCESS) return rc;
    *request = new_request(plan, (int)transpose_type, in, out);
    return *request ? DTFFT_SUCCESS : DTFFT_ERROR_ALLOC_FAILED;
}
dtfft_error_t dtfft_transpose_end(dtfft_plan_t plan, dtfft_request_t request) {
    PLAN_OR_RETURN(plan);
    return end_request(plan, request, true);
}
dtfft_error_t dtfft_reshape(dtfft_plan_t plan, void* in, void* out, dtfft_reshape_t reshape_type, void* aux) {
    PLAN_OR_RETURN(plan);
    if (!in || !out) return DTFFT_ERROR_INVALID_USAGE;
    return E(P(plan)->reshape(in, out, (int)reshape_type, aux));
}
dtfft_error_t dtfft_reshape_start(dtfft_plan_t plan, void* in, void* out, dtfft_reshape_t reshape_type, void* aux,
                                  dtfft_request_t* request) {
    if (!request) return DTFFT_ERROR_INVALID_USAGE;
    *request = nullptr;
    dtfft_error_t rc = dtfft_reshape(plan, in, out, reshape_type, aux);
    if (rc != DTFFT_SUCCESS) return rc;
    *request = new_request(plan, (int)reshape_type, in, out);
    return *request ? DTFFT_SUCCESS : DTFFT_ERROR_ALLOC_FAILED;
}
dtfft_error_t dtfft_reshape_end(dtfft_plan_t plan, dtfft_request_t request) {
    PLAN_OR_RETURN(plan);
    return end_request(plan, request, false);
}

dtfft_error_t dtfft_destroy(dtfft_plan_t* plan) {  // dtfft_api_c.c:190-198: frees and NULLs the handle
    if (!plan || !*plan) return DTFFT_ERROR_PLAN_NOT_CREATED;
    Plan* p = P(*plan);
    drop_requests_of(*plan);
    p->destroy();
    delete p;
    *plan = nullptr;
    return DTFFT_SUCCESS;
}

dtfft_error_t dtfft_get_local_sizes(dtfft_plan_t plan, int32_t* in_starts, int32_t* in_counts, int32_t* out_starts,
                                    int32_t* out_counts, size_t* alloc_size) {
    PLAN_OR_RETURN(plan);
    if (!in_starts && !in_counts && !out_starts && !out_counts && !alloc_size) return DTFFT_ERROR_INVALID_USAGE;
    return E(P(plan)->get_local_sizes(in_starts, in_counts, out_starts, out_counts, alloc_size));
}
dtfft_error_t dtfft_get_alloc_size(dtfft_plan_t plan, size_t* alloc_size) {
    PLAN_OR_RETURN(plan);
    if (!alloc_size) return DTFFT_ERROR_INVALID_USAGE;
    *alloc_size = P(plan)->alloc_size();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_aux_bytes(dtfft_plan_t plan, size_t* aux_bytes) {
    PLAN_OR_RETURN(plan);
    if (!aux_bytes) return DTFFT_ERROR_INVALID_USAGE;
    *aux_bytes = P(plan)->aux_bytes();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_aux_size(dtfft_plan_t plan, size_t* aux_size) {
    PLAN_OR_RETURN(plan);
    if (!aux_size) return DTFFT_ERROR_INVALID_USAGE;
    *aux_size = P(plan)->aux_bytes() / P(plan)->element_size();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_aux_bytes_reshape(dtfft_plan_t plan, size_t* aux_bytes) {
    PLAN_OR_RETURN(plan);
    if (!aux_bytes) return DTFFT_ERROR_INVALID_USAGE;
    if (!P(plan)->reshape_enabled()) return DTFFT_ERROR_RESHAPE_NOT_SUPPORTED;
    *aux_bytes = P(plan)->aux_bytes_reshape();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_aux_size_reshape(dtfft_plan_t plan, size_t* aux_size) {
    size_t b = 0;
    dtfft_error_t rc = dtfft_get_aux_bytes_reshape(plan, aux_size ? &b : nullptr);
    if (rc == DTFFT_SUCCESS) *aux_size = b / P(plan)->element_size();
    return rc;
}
dtfft_error_t dtfft_get_aux_bytes_transpose(dtfft_plan_t plan, size_t* aux_bytes) {
    PLAN_OR_RETURN(plan);
    if (!aux_bytes) return DTFFT_ERROR_INVALID_USAGE;
    *aux_bytes = P(plan)->aux_bytes_transpose();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_aux_size_transpose(dtfft_plan_t plan, size_t* aux_size) {
    size_t b = 0;
    dtfft_error_t rc = dtfft_get_aux_bytes_transpose(plan, aux_size ? &b : nullptr);
    if (rc == DTFFT_SUCCESS) *aux_size = b / P(plan)->element_size();
    return rc;
}
dtfft_error_t dtfft_get_pencil(dtfft_plan_t plan, dtfft_layout_t layout, dtfft_pencil_t* pencil) {
    PLAN_OR_RETURN(plan);
    if (!pencil) return DTFFT_ERROR_INVALID_USAGE;
    return E(P(plan)->get_pencil((int)layout, pencil));
}
dtfft_error_t dtfft_get_element_size(dtfft_plan_t plan, size_t* element_size) {
    PLAN_OR_RETURN(plan);
    if (!element_size) return DTFFT_ERROR_INVALID_USAGE;
    *element_size = P(plan)->element_size();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_alloc_bytes(dtfft_plan_t plan, size_t* alloc_bytes) {
    PLAN_OR_RETURN(plan);
    if (!alloc_bytes) return DTFFT_ERROR_INVALID_USAGE;
    *alloc_bytes = P(plan)->alloc_bytes();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_mem_alloc(dtfft_plan_t plan, size_t alloc_bytes, void** ptr) {
    PLAN_OR_RETURN(plan);
    return E(P(plan)->mem_alloc(alloc_bytes, ptr));
}
dtfft_error_t dtfft_mem_free(dtfft_plan_t plan, void* ptr) {
    PLAN_OR_RETURN(plan);
    if (!ptr) return DTFFT_ERROR_INVALID_USAGE;
    return E(P(plan)->mem_free(ptr));
}
dtfft_error_t dtfft_report(dtfft_plan_t plan) {
    PLAN_OR_RETURN(plan);
    return E(P(plan)->report());
}
dtfft_error_t dtfft_get_z_slab_enabled(dtfft_plan_t plan, bool* v) {
    PLAN_OR_RETURN(plan);
    if (!v) return DTFFT_ERROR_INVALID_USAGE;
    *v = P(plan)->z_slab();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_y_slab_enabled(dtfft_plan_t plan, bool* v) {
    PLAN_OR_RETURN(plan);
    if (!v) return DTFFT_ERROR_INVALID_USAGE;
    *v = P(plan)->y_slab();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_executor(dtfft_plan_t plan, dtfft_executor_t* executor) {
    PLAN_OR_RETURN(plan);
    if (!executor) return DTFFT_ERROR_INVALID_USAGE;
    *executor = (dtfft_executor_t)P(plan)->executor();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_precision(dtfft_plan_t plan, dtfft_precision_t* precision) {
    PLAN_OR_RETURN(plan);
    if (!precision) return DTFFT_ERROR_INVALID_USAGE;
    *precision = (dtfft_precision_t)P(plan)->precision();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_dims(dtfft_plan_t plan, int8_t* ndims, const int32_t* dims[]) {
    PLAN_OR_RETURN(plan);
    if (!ndims && !dims) return DTFFT_ERROR_INVALID_USAGE;
    if (ndims) *ndims = (int8_t)P(plan)->ndims();
    if (dims) *dims = P(plan)->dims();  // owned by the plan (include/dtfft.h:896)
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_grid_dims(dtfft_plan_t plan, int8_t* ndims, const int32_t* grid_dims[]) {
    PLAN_OR_RETURN(plan);
    if (!ndims && !grid_dims) return DTFFT_ERROR_INVALID_USAGE;
    if (ndims) *ndims = (int8_t)P(plan)->ndims();
    if (grid_dims) *grid_dims = P(plan)->grid_dims();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_stream(dtfft_plan_t plan, dtfft_stream_t* stream) {
    PLAN_OR_RETURN(plan);
    if (!stream) return DTFFT_ERROR_INVALID_USAGE;
    *stream = P(plan)->stream();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_platform(dtfft_plan_t plan, dtfft_platform_t* platform) {
    PLAN_OR_RETURN(plan);
    if (!platform) return DTFFT_ERROR_INVALID_USAGE;
    *platform = DTFFT_PLATFORM_CUDA;
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_backend(dtfft_plan_t plan, dtfft_backend_t* backend) {
    PLAN_OR_RETURN(plan);
    if (!backend) return DTFFT_ERROR_INVALID_USAGE;
    *backend = (dtfft_backend_t)P(plan)->backend();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_reshape_backend(dtfft_plan_t plan, dtfft_backend_t* backend) {
    PLAN_OR_RETURN(plan);
    if (!backend) return DTFFT_ERROR_INVALID_USAGE;
    if (!P(plan)->reshape_enabled()) return DTFFT_ERROR_RESHAPE_NOT_SUPPORTED;
    *backend = (dtfft_backend_t)P(plan)->reshape_backend();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfft_get_backend_pipelined(const dtfft_backend_t backend, bool* is_pipe) {
    if (!is_pipe) return DTFFT_ERROR_INVALID_USAGE;
    switch (backend) {
        case DTFFT_BACKEND_MPI_P2P_PIPELINED:
        case DTFFT_BACKEND_NCCL_PIPELINED:
        case DTFFT_BACKEND_CUFFTMP_PIPELINED:
        case DTFFT_BACKEND_MPI_RMA_PIPELINED: *is_pipe = true; break;
        default: *is_pipe = false;
    }
    return DTFFT_SUCCESS;
}

const char* dtfft_get_error_string(dtfft_error_t e) {
    switch ((int)e) {
        case DTFFT_SUCCESS: return "DTFFT_SUCCESS";
        case DTFFT_ERROR_MPI_FINALIZED: return "communicator is not usable (MPI finalized or never initialised)";
        case DTFFT_ERROR_PLAN_NOT_CREATED: return "plan has not been created";
        case DTFFT_ERROR_INVALID_TRANSPOSE_TYPE: return "invalid transpose_type";
        case DTFFT_ERROR_INVALID_N_DIMENSIONS: return "number of dimensions must be 2 or 3";
        case DTFFT_ERROR_INVALID_DIMENSION_SIZE: return "a dimension size is <= 0";
        case DTFFT_ERROR_INVALID_COMM_TYPE: return "invalid communicator type";
        case DTFFT_ERROR_INVALID_PRECISION: return "invalid precision";
        case DTFFT_ERROR_INVALID_EFFORT: return "invalid effort";
        case DTFFT_ERROR_INVALID_EXECUTOR: return "invalid executor";
        case DTFFT_ERROR_INVALID_COMM_DIMS: return "process grid has more dimensions than the plan";
        case DTFFT_ERROR_INVALID_COMM_FAST_DIM: return "the fastest dimension must not be distributed";
        case DTFFT_ERROR_MISSING_R2R_KINDS: return "R2R plan with an executor needs `kinds`";
        case DTFFT_ERROR_INVALID_R2R_KINDS: return "invalid value in `kinds`";
        case DTFFT_ERROR_R2C_TRANSPOSE_PLAN: return "transpose-only plans are not available for R2C; use C2C or R2R";
        case DTFFT_ERROR_INPLACE_TRANSPOSE: return "in-place transpose is not supported";
        case DTFFT_ERROR_INVALID_AUX: return "invalid aux buffer";
        case DTFFT_ERROR_INVALID_LAYOUT: return "invalid layout";
        case DTFFT_ERROR_INVALID_USAGE: return "invalid usage (NULL argument?)";
        case DTFFT_ERROR_PLAN_IS_CREATED: return "plan is already created";
        case DTFFT_ERROR_ALLOC_FAILED: return "memory allocation failed";
        case DTFFT_ERROR_FREE_FAILED: return "memory free failed";
        case DTFFT_ERROR_INVALID_ALLOC_BYTES: return "invalid alloc_bytes";
        case DTFFT_ERROR_PENCIL_ARRAYS_SIZE_MISMATCH: return "pencil starts/counts have different sizes";
        case DTFFT_ERROR_PENCIL_ARRAYS_INVALID_SIZES: return "pencil must have 2 or 3 dimensions";
        case DTFFT_ERROR_PENCIL_INVALID_COUNTS: return "pencil counts < 0";
        case DTFFT_ERROR_PENCIL_INVALID_STARTS: return "pencil starts < 0";
        case DTFFT_ERROR_PENCIL_SHAPE_MISMATCH: return "pencils with equal starts have different shapes";
        case DTFFT_ERROR_PENCIL_OVERLAP: return "pencils overlap";
        case DTFFT_ERROR_PENCIL_NOT_CONTINUOUS: return "pencils do not tile the global domain";
        case DTFFT_ERROR_PENCIL_NOT_INITIALIZED: return "pencil is not initialised";
        case DTFFT_ERROR_INVALID_MEASURE_WARMUP_ITERS: return "invalid n_measure_warmup_iters";
        case DTFFT_ERROR_INVALID_MEASURE_ITERS: return "invalid n_measure_iters";
        case DTFFT_ERROR_INVALID_REQUEST: return "invalid request";
        case DTFFT_ERROR_TRANSPOSE_ACTIVE: return "a transposition is still active";
        case DTFFT_ERROR_TRANSPOSE_NOT_ACTIVE: return "no active transposition";
        case DTFFT_ERROR_INVALID_RESHAPE_TYPE: return "invalid reshape_type";
        case DTFFT_ERROR_RESHAPE_ACTIVE: return "a reshape is still active";
        case DTFFT_ERROR_RESHAPE_NOT_ACTIVE: return "no active reshape";
        case DTFFT_ERROR_INPLACE_RESHAPE: return "in-place reshape is not supported";
        case DTFFT_ERROR_INVALID_EXECUTE_TYPE: return "invalid execute_type";
        case DTFFT_ERROR_RESHAPE_NOT_SUPPORTED: return "plan was not created from bricks: reshape unavailable";
        case DTFFT_ERROR_R2C_EXECUTE_CALLED: return "execute called on an R2C transpose plan";
        case DTFFT_ERROR_INVALID_CART_COMM: return "invalid process grid for the brick decomposition";
        case DTFFT_ERROR_INVALID_TRANSPOSE_MODE: return "invalid transpose_mode";
        case DTFFT_ERROR_INVALID_ACCESS_MODE: return "invalid access_mode";
        case DTFFT_ERROR_R2R_FFT_NOT_SUPPORTED: return "the executor has no R2R transforms";
        case DTFFT_ERROR_GPU_INVALID_STREAM: return "invalid CUDA stream";
        case DTFFT_ERROR_INVALID_BACKEND: return "invalid backend";
        case DTFFT_ERROR_GPU_NOT_SET: return "ranks of one host must use distinct GPUs";
        case DTFFT_ERROR_BACKENDS_DISABLED: return "every usable backend is disabled";
        case DTFFT_ERROR_NOT_DEVICE_PTR: return "a buffer is not a device pointer";
        case DTFFT_ERROR_INVALID_PLATFORM: return "invalid platform (this library is CUDA only)";
        case DTFFT_ERROR_INVALID_PLATFORM_EXECUTOR: return "executor not available on this platform";
        case DTFFT_ERROR_INVALID_PLATFORM_BACKEND: return "backend not available on this platform";
        case DTFFTB_ERROR_NOT_REGISTERED: return "NVLINK_FUSED: `out` must be a registered (dtfft_mem_alloc) buffer";
        case DTFFTB_ERROR_COMM: return "host allgather callback failed";
        case DTFFTB_ERROR_PEER_TIMEOUT: return "NVLINK_FUSED: a group member missed a device barrier (DTFFTB_PEER_TIMEOUT_MS); the plan is dead";
        case DTFFTB_ERROR_INTERNAL: return "internal error";
        default:
            if ((int)e <= DTFFTB_ERROR_NCCL_BASE && (int)e > DTFFTB_ERROR_INTERNAL) return "NCCL error";
            if ((int)e <= DTFFTB_ERROR_CUDA_BASE && (int)e > DTFFTB_ERROR_NCCL_BASE) return "CUDA / cuFFT error";
            return "unknown error";
    }
}
const char* dtfft_get_precision_string(dtfft_precision_t p) {
    return p == DTFFT_SINGLE ? "Single" : p == DTFFT_DOUBLE ? "Double" : "Unknown precision";
}
const char* dtfft_get_executor_string(dtfft_executor_t e) {
    switch (e) {
        case DTFFT_EXECUTOR_NONE: return "None";
        case DTFFT_EXECUTOR_FFTW3: return "FFTW3";
        case DTFFT_EXECUTOR_MKL: return "MKL";
        case DTFFT_EXECUTOR_CUFFT: return "CUFFT";
        case DTFFT_EXECUTOR_VKFFT: return "VKFFT";
        default: return "Unknown executor";
    }
}
const char* dtfft_get_backend_string(dtfft_backend_t b) {
    switch (b) {
        case DTFFT_BACKEND_MPI_DATATYPE: return "MPI_DATATYPE";
        case DTFFT_BACKEND_MPI_P2P: return "MPI_P2P";
        case DTFFT_BACKEND_MPI_A2A: return "MPI_A2A";
        case DTFFT_BACKEND_NCCL: return "NCCL";
        case DTFFT_BACKEND_CUFFTMP: return "CUFFTMP";
        case DTFFT_BACKEND_MPI_P2P_PIPELINED: return "MPI_P2P_PIPELINED";
        case DTFFT_BACKEND_NCCL_PIPELINED: return "NCCL_PIPELINED";
        case DTFFT_BACKEND_CUFFTMP_PIPELINED: return "CUFFTMP_PIPELINED";
        case DTFFT_BACKEND_MPI_RMA: return "MPI_RMA";
        case DTFFT_BACKEND_MPI_RMA_PIPELINED: return "MPI_RMA_PIPELINED";
        case DTFFT_BACKEND_MPI_P2P_SCHEDULED: return "MPI_P2P_SCHEDULED";
        case DTFFT_BACKEND_MPI_P2P_FUSED: return "MPI_P2P_FUSED";
        case DTFFT_BACKEND_MPI_RMA_FUSED: return "MPI_RMA_FUSED";
        case DTFFT_BACKEND_MPI_P2P_COMPRESSED: return "MPI_P2P_COMPRESSED";
        case DTFFT_BACKEND_MPI_RMA_COMPRESSED: return "MPI_RMA_COMPRESSED";
        case DTFFT_BACKEND_ADAPTIVE: return "ADAPTIVE";
        case DTFFT_BACKEND_NCCL_COMPRESSED: return "NCCL_COMPRESSED";
        case DTFFT_BACKEND_NVLINK_FUSED: return "NVLINK_FUSED";
        case DTFFT_BACKEND_NONE: return "NONE";
        default: return "Unknown backend";
    }
}

dtfft_error_t dtfft_create_config(dtfft_config_t* c) {  // src/dtfft_config.F90:644-669
    if (!c) return DTFFT_ERROR_INVALID_USAGE;
    dtfftb::Config d;
    c->enable_log = d.enable_log, c->enable_z_slab = d.enable_z_slab, c->enable_y_slab = d.enable_y_slab;
    c->n_measure_warmup_iters = d.n_measure_warmup_iters, c->n_measure_iters = d.n_measure_iters;
    c->platform = DTFFT_PLATFORM_CUDA;
    c->stream = nullptr;
    c->backend = DTFFT_BACKEND_NONE, c->reshape_backend = DTFFT_BACKEND_NONE;
    c->enable_datatype_backend = d.enable_datatype_backend, c->enable_mpi_backends = d.enable_mpi_backends;
    c->enable_pipelined_backends = d.enable_pipelined_backends, c->enable_rma_backends = d.enable_rma_backends;
    c->enable_fused_backends = d.enable_fused_backends, c->enable_nccl_backends = d.enable_nccl_backends;
    c->enable_nvshmem_backends = d.enable_nvshmem_backends, c->enable_kernel_autotune = d.enable_kernel_autotune;
    c->enable_fourier_reshape = d.enable_fourier_reshape;
    c->transpose_mode = DTFFT_TRANSPOSE_MODE_PACK, c->access_mode = DTFFT_ACCESS_MODE_WRITE;
    return DTFFT_SUCCESS;
}

dtfft_error_t dtfft_set_config(const dtfft_config_t* c) {  // src/dtfft_config.F90:677-764
    if (!c) return DTFFT_ERROR_INVALID_USAGE;
    if (c->n_measure_warmup_iters < 0) return DTFFT_ERROR_INVALID_MEASURE_WARMUP_ITERS;
    if (c->n_measure_iters < 1) return DTFFT_ERROR_INVALID_MEASURE_ITERS;
    if (c->platform != DTFFT_PLATFORM_CUDA) return DTFFT_ERROR_INVALID_PLATFORM;
    auto ok_backend = [](int b) {
        return b == DTFFT_BACKEND_NONE || b == DTFFT_BACKEND_NCCL || b == DTFFT_BACKEND_NCCL_PIPELINED ||
               b == DTFFT_BACKEND_NVLINK_FUSED;
    };
    auto known_backend = [](int b) { return b == DTFFT_BACKEND_NONE || (b >= 21 && b <= 38); };
    if (!known_backend(c->backend) || !known_backend(c->reshape_backend)) return DTFFT_ERROR_INVALID_BACKEND;
    if (!ok_backend(c->backend) || !ok_backend(c->reshape_backend)) return DTFFT_ERROR_INVALID_PLATFORM_BACKEND;
    if (c->transpose_mode != DTFFT_TRANSPOSE_MODE_PACK && c->transpose_mode != DTFFT_TRANSPOSE_MODE_UNPACK)
        return DTFFT_ERROR_INVALID_TRANSPOSE_MODE;
    if (c->access_mode != DTFFT_ACCESS_MODE_WRITE && c->access_mode != DTFFT_ACCESS_MODE_READ)
        return DTFFT_ERROR_INVALID_ACCESS_MODE;
    if (c->stream) {
        cudaError_t ce = cudaStreamQuery(static_cast<cudaStream_t>(c->stream));
        if (ce != cudaSuccess && ce != cudaErrorNotReady) {
            cudaGetLastError();
            return DTFFT_ERROR_GPU_INVALID_STREAM;
        }
    }
    dtfftb::Config& g = dtfftb::global_config();
    g.enable_log = c->enable_log, g.enable_z_slab = c->enable_z_slab, g.enable_y_slab = c->enable_y_slab;
    g.n_measure_warmup_iters = c->n_measure_warmup_iters, g.n_measure_iters = c->n_measure_iters;
    g.platform = c->platform, g.stream = c->stream;
    g.backend = c->backend, g.reshape_backend = c->reshape_backend;
    g.enable_datatype_backend = c->enable_datatype_backend, g.enable_mpi_backends = c->enable_mpi_backends;
    g.enable_pipelined_backends = c->enable_pipelined_backends, g.enable_rma_backends = c->enable_rma_backends;
    g.enable_fused_backends = c->enable_fused_backends, g.enable_nccl_backends = c->enable_nccl_backends;
    g.enable_nvshmem_backends = c->enable_nvshmem_backends, g.enable_kernel_autotune = c->enable_kernel_autotune;
    g.enable_fourier_reshape = c->enable_fourier_reshape;
    g.transpose_mode = c->transpose_mode, g.access_mode = c->access_mode;
    return DTFFT_SUCCESS;
}

dtfft_error_t dtfftb_plan_register_buffer(dtfft_plan_t plan, void* ptr, size_t bytes) {
    PLAN_OR_RETURN(plan);
    if (!ptr || !bytes) return DTFFT_ERROR_INVALID_USAGE;
    return E(P(plan)->register_buffer(ptr, bytes));
}
dtfft_error_t dtfftb_plan_unregister_buffer(dtfft_plan_t plan, void* ptr) {
    PLAN_OR_RETURN(plan);
    return E(P(plan)->unregister_buffer(ptr));
}
dtfft_error_t dtfftb_plan_get_stats(dtfft_plan_t plan, int64_t* kernel_launches, int64_t* local_bytes,
                                    int64_t* remote_bytes) {
    PLAN_OR_RETURN(plan);
    int64_t a, b, c;
    P(plan)->last_stats(&a, &b, &c);
    if (kernel_launches) *kernel_launches = a;
    if (local_bytes) *local_bytes = b;
    if (remote_bytes) *remote_bytes = c;
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfftb_plan_get_exchange_form(dtfft_plan_t plan, int transpose_type, int* form, int* n_slices) {
    PLAN_OR_RETURN(plan);
    int f = 0, n = 0;
    int rc = P(plan)->exchange_form(transpose_type, &f, &n);
    if (rc) return E(rc);
    if (form) *form = f;
    if (n_slices) *n_slices = n;
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfftb_plan_set_overlap(dtfft_plan_t plan, int nchunks, int exchange_ctas) {
    PLAN_OR_RETURN(plan);
    P(plan)->set_overlap(nchunks < 1 ? 1 : nchunks, exchange_ctas < 0 ? 0 : exchange_ctas);
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfftb_plan_get_overlap(dtfft_plan_t plan, int* nchunks) {
    PLAN_OR_RETURN(plan);
    if (!nchunks) return DTFFT_ERROR_INVALID_USAGE;
    *nchunks = P(plan)->overlap_chunks();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfftb_plan_set_graphs(dtfft_plan_t plan, int enable) {
    PLAN_OR_RETURN(plan);
    P(plan)->set_graphs(enable != 0);
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfftb_plan_get_graph_replays(dtfft_plan_t plan, int64_t* n_replays) {
    PLAN_OR_RETURN(plan);
    if (!n_replays) return DTFFT_ERROR_INVALID_USAGE;
    *n_replays = P(plan)->graph_replays();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfftb_plan_get_fallbacks(dtfft_plan_t plan, int64_t* n_fallbacks) {
    PLAN_OR_RETURN(plan);
    if (!n_fallbacks) return DTFFT_ERROR_INVALID_USAGE;
    *n_fallbacks = P(plan)->fallbacks();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfftb_plan_get_overlapped_stages(dtfft_plan_t plan, int64_t* n_stages) {
    PLAN_OR_RETURN(plan);
    if (!n_stages) return DTFFT_ERROR_INVALID_USAGE;
    *n_stages = P(plan)->overlapped_stages();
    return DTFFT_SUCCESS;
}
dtfft_error_t dtfftb_plan_create_dry(int kind, int8_t ndims, const int32_t* dims, const dtfft_pencil_t* pencil,
                                     dtfft_comm_t comm, dtfft_precision_t precision, dtfft_executor_t executor,
                                     dtfft_plan_t* plan) {
    if (!plan) return DTFFT_ERROR_INVALID_USAGE;
    *plan = nullptr;
    if (!dims && !pencil) return DTFFT_ERROR_INVALID_USAGE;
    if (kind < 0 || kind > 2) return DTFFT_ERROR_INVALID_USAGE;
    Plan* p = new (std::nothrow) Plan;
    if (!p) return DTFFT_ERROR_ALLOC_FAILED;
    int rc = p->create((dtfftb::PlanKind)kind, ndims, dims, pencil, nullptr, comm, (int)precision, DTFFT_ESTIMATE,
                       (int)executor, true);
    if (rc) {
        delete p;
        return E(rc);
    }
    *plan = p;
    return DTFFT_SUCCESS;
}

dtfft_error_t dtfftb_plan_dry_set_grid(dtfft_plan_t plan, int32_t g1, int32_t g2) {
    PLAN_OR_RETURN(plan);
    return E(P(plan)->dry_set_grid(g1, g2));
}

int32_t dtfftb_grid_candidates(const int32_t* dims, int32_t comm_size, int32_t cap, int32_t* grids) {
    if (!dims || comm_size < 1) return -1;
    const auto g = dtfftb::grid_candidates(dims, comm_size);
    for (size_t i = 0; i < g.size() && (int32_t)i < cap && grids; ++i) grids[2 * i] = g[i].first, grids[2 * i + 1] = g[i].second;
    return (int32_t)g.size();
}

dtfft_error_t dtfftb_plan_describe_exchange(dtfft_plan_t plan, int type, int32_t cap, int32_t* n_members,
                                            int32_t* my_index, int32_t* members, int32_t* kernels, int32_t* send_nd,
                                            int32_t* recv_nd, int64_t* counts_displs, int64_t* fused_boxes,
                                            int32_t* fused_transposing) {
    PLAN_OR_RETURN(plan);
    if (!n_members) return DTFFT_ERROR_INVALID_USAGE;
    Plan::ExchangeDescription d;
    int rc = P(plan)->describe_exchange(type, &d);
    if (rc) return E(rc);
    const int np = (int)d.members.size();
    *n_members = np;
    if (my_index) *my_index = d.me;
    if (np > cap) return DTFFT_SUCCESS;  // caller re-queries with a larger capacity
    for (int i = 0; i < np; ++i) {
        if (members) members[i] = d.members[(size_t)i];
        if (fused_boxes) {
            const dtfftb::Box& b = d.fused[(size_t)i];
            int64_t* o = fused_boxes + 10 * i;
            o[0] = b.empty() ? 0 : b.n0, o[1] = b.n1, o[2] = b.n2, o[3] = b.in_off, o[4] = b.out_off;
            o[5] = b.is1, o[6] = b.is2, o[7] = b.os0, o[8] = b.os1, o[9] = b.os2;
        }
    }
    if (fused_transposing) *fused_transposing = d.fused_transposing ? 1 : 0;
    if (kernels) kernels[0] = d.geo.pack_kernel, kernels[1] = d.geo.unpack_kernel;
    const bool has = !d.geo.send_nd.empty();
    for (int i = 0; i < np && has; ++i) {
        for (int j = 0; j < 5; ++j) {
            if (send_nd) send_nd[5 * i + j] = d.geo.send_nd[(size_t)(5 * i + j)];
            if (recv_nd) recv_nd[5 * i + j] = d.geo.recv_nd[(size_t)(5 * i + j)];
        }
        if (counts_displs) {
            counts_displs[0 * np + i] = d.geo.send_counts[(size_t)i];
            counts_displs[1 * np + i] = d.geo.send_displs[(size_t)i];
            counts_displs[2 * np + i] = d.geo.recv_counts[(size_t)i];
            counts_displs[3 * np + i] = d.geo.recv_displs[(size_t)i];
        }
    }
    return DTFFT_SUCCESS;
}

dtfft_error_t dtfftb_plan_describe_reshape(dtfft_plan_t plan, int reshape_type, int32_t cap, int32_t* n_members,
                                           int32_t* my_index, int32_t* members, int64_t* pack_boxes,
                                           int64_t* unpack_boxes, int64_t* counts_displs, int32_t* flags) {
    PLAN_OR_RETURN(plan);
    if (!n_members) return DTFFT_ERROR_INVALID_USAGE;
    std::vector<int> mem;
    int me = 0;
    dtfftb::ReshapeGeometry g;
    int rc = P(plan)->describe_reshape(reshape_type, &mem, &me, &g);
    if (rc) return E(rc);
    const int np = (int)mem.size();
    *n_members = np;
    if (my_index) *my_index = me;
    if (flags) flags[0] = g.is_pack_free ? 1 : 0, flags[1] = g.is_unpack_free ? 1 : 0, flags[2] = g.reshape_strat;
    if (np > cap) return DTFFT_SUCCESS;  // caller re-queries with a larger capacity
    auto put = [](int64_t* o, const dtfftb::Box& b) {
        o[0] = b.empty() ? 0 : b.n0, o[1] = b.n1, o[2] = b.n2, o[3] = b.in_off, o[4] = b.out_off;
        o[5] = b.is1, o[6] = b.is2, o[7] = b.os0, o[8] = b.os1, o[9] = b.os2;
    };
    for (int i = 0; i < np; ++i) {
        if (members) members[i] = mem[(size_t)i];
        if (pack_boxes) put(pack_boxes + 10 * i, g.pack_boxes[(size_t)i]);
        if (unpack_boxes) put(unpack_boxes + 10 * i, g.unpack_boxes[(size_t)i]);
        if (counts_displs) {
            counts_displs[0 * np + i] = g.send_counts[(size_t)i];
            counts_displs[1 * np + i] = g.send_displs[(size_t)i];
            counts_displs[2 * np + i] = g.recv_counts[(size_t)i];
            counts_displs[3 * np + i] = g.recv_displs[(size_t)i];
        }
    }
    return DTFFT_SUCCESS;
}

dtfft_error_t dtfftb_plan_describe_chunk(dtfft_plan_t plan, int transpose_type, int32_t k, int32_t nchunks, int32_t cap,
                                         int32_t* n_members, int64_t* boxes, int64_t* chunk_offset) {
    PLAN_OR_RETURN(plan);
    if (!n_members) return DTFFT_ERROR_INVALID_USAGE;
    std::vector<int> mem;
    std::vector<dtfftb::Box> bx;
    long long off = 0;
    int rc = P(plan)->describe_chunk(transpose_type, k, nchunks, &mem, &bx, &off);
    if (rc) return E(rc);
    *n_members = (int32_t)mem.size();
    if (chunk_offset) *chunk_offset = off;
    if (!boxes || (int)mem.size() > cap) return DTFFT_SUCCESS;
    for (size_t i = 0; i < bx.size(); ++i) {
        const dtfftb::Box& b = bx[i];
        int64_t* o = boxes + 10 * i;
        o[0] = b.empty() ? 0 : b.n0, o[1] = b.n1, o[2] = b.n2, o[3] = b.in_off, o[4] = b.out_off;
        o[5] = b.is1, o[6] = b.is2, o[7] = b.os0, o[8] = b.os1, o[9] = b.os2;
    }
    return DTFFT_SUCCESS;
}

dtfft_error_t dtfftb_plan_describe_dma(dtfft_plan_t plan, int ttype, int32_t cap_members, int32_t cap_entries,
                                       int32_t* n_members, int32_t* me, int32_t* members, int32_t* n_entries, int64_t* rows) {
    PLAN_OR_RETURN(plan);
    if (!n_members || !n_entries) return DTFFT_ERROR_INVALID_USAGE;
    std::vector<int> mem;
    int my = 0;
    std::vector<dtfftb::Plan::DmaEntry> ent;
    int rc = P(plan)->describe_dma(ttype, &mem, &my, &ent);
    if (rc) return E(rc);
    *n_members = (int32_t)mem.size();
    *n_entries = (int32_t)ent.size();
    if (me) *me = my;
    if (!rows || !members || (int)mem.size() > cap_members || (int)ent.size() > cap_entries) return DTFFT_SUCCESS;
    for (size_t i = 0; i < mem.size(); ++i) members[i] = mem[i];
    for (size_t i = 0; i < ent.size(); ++i) {
        int64_t* o = rows + 30 * i;
        const dtfftb::Box& b = ent[i].blk.pack;
        o[0] = b.empty() ? 0 : b.n0, o[1] = b.n1, o[2] = b.n2, o[3] = b.in_off, o[4] = b.out_off;
        o[5] = b.is1, o[6] = b.is2, o[7] = b.os0, o[8] = b.os1, o[9] = b.os2;
        o[10] = ent[i].blk.run, o[11] = ent[i].blk.rows, o[12] = ent[i].blk.planes, o[13] = ent[i].blk.dst_off;
        o[14] = ent[i].blk.dst_pitch, o[15] = ent[i].blk.dst_plane_rows, o[16] = ent[i].blk.ok ? 1 : 0;
        const dtfftb::Box& f = ent[i].fused;
        o[17] = f.empty() ? 0 : f.n0, o[18] = f.n1, o[19] = f.n2, o[20] = f.in_off, o[21] = f.out_off;
        o[22] = f.is1, o[23] = f.is2, o[24] = f.os0, o[25] = f.os1, o[26] = f.os2;
        o[27] = ent[i].member, o[28] = ent[i].sub, o[29] = ent[i].nsub;
    }
    return DTFFT_SUCCESS;
}

dtfft_error_t dtfftb_plan_describe_peer_piece(dtfft_plan_t plan, int t_local, int t_exchange, int side, int32_t peer,
                                              int32_t sub, int32_t* nsub, int64_t* box) {
    PLAN_OR_RETURN(plan);
    if (!box) return DTFFT_ERROR_INVALID_USAGE;
    dtfftb::Box b;
    int ns = 1;
    int rc = P(plan)->describe_peer_piece(t_local, t_exchange, side, peer, sub, &ns, &b);
    if (nsub) *nsub = ns;
    if (rc) return E(rc);
    box[0] = b.empty() ? 0 : b.n0, box[1] = b.n1, box[2] = b.n2, box[3] = b.in_off, box[4] = b.out_off;
    box[5] = b.is1, box[6] = b.is2, box[7] = b.os0, box[8] = b.os1, box[9] = b.os2;
    return DTFFT_SUCCESS;
}

int dtfftb_plan_peer_error(dtfft_plan_t plan) {
    if (!plan) return 0;
    return P(plan)->peer_error();
}

}  // extern "C"
