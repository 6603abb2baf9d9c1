/* dtfft_b200.hpp -- header-only C++17 mirror of the reference's C++ API (namespace dtfft, classes
 * Plan / PlanC2C / PlanR2C / PlanR2R / Pencil / Config / Exception; reference: include/dtfft.hpp:62-2238,
 * implemented there in src/interfaces/api/dtfft_api_cxx.cpp purely over the C ABI).  Here the same
 * surface sits over include/dtfft_b200_api.h, so a C++ caller of the reference changes two things:
 *   #include <dtfft.hpp>  ->  #include <dtfft_b200.hpp>
 *   MPI_Comm arguments    ->  dtfft_comm_t (a dtfftb_comm_t*: rank, size, one allgather callback;
 *                             nullptr = single rank; include/dtfft_b200_mpi.h wraps an MPI_Comm)
 * Every operation exists twice, as in the reference: a noexcept form returning dtfft::Error and a
 * throwing convenience form returning the value (dtfft::Exception carries the error code).
 * Extensions of this library (NVLink registration, stage overlap, CUDA graphs) are the methods
 * whose names do not exist upstream: register_buffer, set_overlap, set_graphs, get_stats. */
#ifndef DTFFT_B200_HPP
#define DTFFT_B200_HPP

#include <complex>
#include <cstdint>
#include <exception>
#include <string>
#include <utility>
#include <vector>

#include "dtfft_b200_api.h"

namespace dtfft {

struct Version {
    static constexpr int32_t MAJOR = DTFFT_VERSION_MAJOR, MINOR = DTFFT_VERSION_MINOR, PATCH = DTFFT_VERSION_PATCH;
    static constexpr int32_t CODE = DTFFT_VERSION_CODE;
    static int32_t get() noexcept { return dtfft_get_version(); }
    static constexpr int32_t get(int32_t major, int32_t minor, int32_t patch) noexcept {
        return DTFFT_VERSION(major, minor, patch);
    }
};

/* Same numeric values as dtfft_error_t; only the codes a CUDA build can return are named, any other
 * value still converts (static_cast) and prints through get_error_string. */
enum class Error : int {
    SUCCESS = DTFFT_SUCCESS,
    MPI_FINALIZED = DTFFT_ERROR_MPI_FINALIZED,
    PLAN_NOT_CREATED = DTFFT_ERROR_PLAN_NOT_CREATED,
    INVALID_TRANSPOSE_TYPE = DTFFT_ERROR_INVALID_TRANSPOSE_TYPE,
    INVALID_N_DIMENSIONS = DTFFT_ERROR_INVALID_N_DIMENSIONS,
    INVALID_DIMENSION_SIZE = DTFFT_ERROR_INVALID_DIMENSION_SIZE,
    INVALID_COMM_TYPE = DTFFT_ERROR_INVALID_COMM_TYPE,
    INVALID_PRECISION = DTFFT_ERROR_INVALID_PRECISION,
    INVALID_EFFORT = DTFFT_ERROR_INVALID_EFFORT,
    INVALID_EXECUTOR = DTFFT_ERROR_INVALID_EXECUTOR,
    INVALID_COMM_DIMS = DTFFT_ERROR_INVALID_COMM_DIMS,
    INVALID_COMM_FAST_DIM = DTFFT_ERROR_INVALID_COMM_FAST_DIM,
    MISSING_R2R_KINDS = DTFFT_ERROR_MISSING_R2R_KINDS,
    INVALID_R2R_KINDS = DTFFT_ERROR_INVALID_R2R_KINDS,
    R2C_TRANSPOSE_PLAN = DTFFT_ERROR_R2C_TRANSPOSE_PLAN,
    INPLACE_TRANSPOSE = DTFFT_ERROR_INPLACE_TRANSPOSE,
    INVALID_AUX = DTFFT_ERROR_INVALID_AUX,
    INVALID_LAYOUT = DTFFT_ERROR_INVALID_LAYOUT,
    INVALID_USAGE = DTFFT_ERROR_INVALID_USAGE,
    PLAN_IS_CREATED = DTFFT_ERROR_PLAN_IS_CREATED,
    ALLOC_FAILED = DTFFT_ERROR_ALLOC_FAILED,
    FREE_FAILED = DTFFT_ERROR_FREE_FAILED,
    INVALID_ALLOC_BYTES = DTFFT_ERROR_INVALID_ALLOC_BYTES,
    PENCIL_ARRAYS_SIZE_MISMATCH = DTFFT_ERROR_PENCIL_ARRAYS_SIZE_MISMATCH,
    PENCIL_ARRAYS_INVALID_SIZES = DTFFT_ERROR_PENCIL_ARRAYS_INVALID_SIZES,
    PENCIL_INVALID_COUNTS = DTFFT_ERROR_PENCIL_INVALID_COUNTS,
    PENCIL_INVALID_STARTS = DTFFT_ERROR_PENCIL_INVALID_STARTS,
    PENCIL_SHAPE_MISMATCH = DTFFT_ERROR_PENCIL_SHAPE_MISMATCH,
    PENCIL_OVERLAP = DTFFT_ERROR_PENCIL_OVERLAP,
    PENCIL_NOT_CONTINUOUS = DTFFT_ERROR_PENCIL_NOT_CONTINUOUS,
    PENCIL_NOT_INITIALIZED = DTFFT_ERROR_PENCIL_NOT_INITIALIZED,
    INVALID_MEASURE_WARMUP_ITERS = DTFFT_ERROR_INVALID_MEASURE_WARMUP_ITERS,
    INVALID_MEASURE_ITERS = DTFFT_ERROR_INVALID_MEASURE_ITERS,
    INVALID_REQUEST = DTFFT_ERROR_INVALID_REQUEST,
    TRANSPOSE_ACTIVE = DTFFT_ERROR_TRANSPOSE_ACTIVE,
    TRANSPOSE_NOT_ACTIVE = DTFFT_ERROR_TRANSPOSE_NOT_ACTIVE,
    INVALID_RESHAPE_TYPE = DTFFT_ERROR_INVALID_RESHAPE_TYPE,
    RESHAPE_ACTIVE = DTFFT_ERROR_RESHAPE_ACTIVE,
    RESHAPE_NOT_ACTIVE = DTFFT_ERROR_RESHAPE_NOT_ACTIVE,
    INPLACE_RESHAPE = DTFFT_ERROR_INPLACE_RESHAPE,
    INVALID_EXECUTE_TYPE = DTFFT_ERROR_INVALID_EXECUTE_TYPE,
    RESHAPE_NOT_SUPPORTED = DTFFT_ERROR_RESHAPE_NOT_SUPPORTED,
    R2C_EXECUTE_CALLED = DTFFT_ERROR_R2C_EXECUTE_CALLED,
    INVALID_CART_COMM = DTFFT_ERROR_INVALID_CART_COMM,
    INVALID_TRANSPOSE_MODE = DTFFT_ERROR_INVALID_TRANSPOSE_MODE,
    INVALID_ACCESS_MODE = DTFFT_ERROR_INVALID_ACCESS_MODE,
    R2R_FFT_NOT_SUPPORTED = DTFFT_ERROR_R2R_FFT_NOT_SUPPORTED,
    GPU_INVALID_STREAM = DTFFT_ERROR_GPU_INVALID_STREAM,
    INVALID_BACKEND = DTFFT_ERROR_INVALID_BACKEND,
    GPU_NOT_SET = DTFFT_ERROR_GPU_NOT_SET,
    BACKENDS_DISABLED = DTFFT_ERROR_BACKENDS_DISABLED,
    NOT_DEVICE_PTR = DTFFT_ERROR_NOT_DEVICE_PTR,
    INVALID_PLATFORM = DTFFT_ERROR_INVALID_PLATFORM,
    INVALID_PLATFORM_EXECUTOR = DTFFT_ERROR_INVALID_PLATFORM_EXECUTOR,
    INVALID_PLATFORM_BACKEND = DTFFT_ERROR_INVALID_PLATFORM_BACKEND
};

enum class Execute : int { FORWARD = DTFFT_EXECUTE_FORWARD, BACKWARD = DTFFT_EXECUTE_BACKWARD };
enum class Transpose : int {
    X_TO_Y = DTFFT_TRANSPOSE_X_TO_Y, Y_TO_X = DTFFT_TRANSPOSE_Y_TO_X, Y_TO_Z = DTFFT_TRANSPOSE_Y_TO_Z,
    Z_TO_Y = DTFFT_TRANSPOSE_Z_TO_Y, X_TO_Z = DTFFT_TRANSPOSE_X_TO_Z, Z_TO_X = DTFFT_TRANSPOSE_Z_TO_X
};
enum class Reshape : int {
    X_BRICKS_TO_PENCILS = DTFFT_RESHAPE_X_BRICKS_TO_PENCILS, X_PENCILS_TO_BRICKS = DTFFT_RESHAPE_X_PENCILS_TO_BRICKS,
    Z_PENCILS_TO_BRICKS = DTFFT_RESHAPE_Z_PENCILS_TO_BRICKS, Z_BRICKS_TO_PENCILS = DTFFT_RESHAPE_Z_BRICKS_TO_PENCILS,
    Y_BRICKS_TO_PENCILS = DTFFT_RESHAPE_Y_BRICKS_TO_PENCILS, Y_PENCILS_TO_BRICKS = DTFFT_RESHAPE_Y_PENCILS_TO_BRICKS
};
enum class Layout : int {
    X_BRICKS = DTFFT_LAYOUT_X_BRICKS, X_PENCILS = DTFFT_LAYOUT_X_PENCILS,
    X_PENCILS_FOURIER = DTFFT_LAYOUT_X_PENCILS_FOURIER, Y_PENCILS = DTFFT_LAYOUT_Y_PENCILS,
    Z_PENCILS = DTFFT_LAYOUT_Z_PENCILS, Z_BRICKS = DTFFT_LAYOUT_Z_BRICKS
};
enum class Precision : int { SINGLE = DTFFT_SINGLE, DOUBLE = DTFFT_DOUBLE };
enum class Effort : int {
    ESTIMATE = DTFFT_ESTIMATE, MEASURE = DTFFT_MEASURE, PATIENT = DTFFT_PATIENT, EXHAUSTIVE = DTFFT_EXHAUSTIVE
};
enum class Executor : int {
    NONE = DTFFT_EXECUTOR_NONE, FFTW3 = DTFFT_EXECUTOR_FFTW3, MKL = DTFFT_EXECUTOR_MKL,
    CUFFT = DTFFT_EXECUTOR_CUFFT, VKFFT = DTFFT_EXECUTOR_VKFFT
};
enum class R2RKind : int {
    DCT_1 = DTFFT_DCT_1, DCT_2 = DTFFT_DCT_2, DCT_3 = DTFFT_DCT_3, DCT_4 = DTFFT_DCT_4,
    DST_1 = DTFFT_DST_1, DST_2 = DTFFT_DST_2, DST_3 = DTFFT_DST_3, DST_4 = DTFFT_DST_4
};
/* Backends that exist on this path; the MPI / cuFFTMp / compressed values of the reference are
 * accepted by the C enum and rejected at plan creation (Error::INVALID_BACKEND). */
enum class Backend : int {
    NCCL = DTFFT_BACKEND_NCCL, NCCL_PIPELINED = DTFFT_BACKEND_NCCL_PIPELINED,
    NVLINK_FUSED = DTFFT_BACKEND_NVLINK_FUSED, NONE = DTFFT_BACKEND_NONE
};
enum class Platform : int { HOST = DTFFT_PLATFORM_HOST, CUDA = DTFFT_PLATFORM_CUDA };
enum class TransposeMode : int { PACK = DTFFT_TRANSPOSE_MODE_PACK, UNPACK = DTFFT_TRANSPOSE_MODE_UNPACK };
enum class AccessMode : int { WRITE = DTFFT_ACCESS_MODE_WRITE, READ = DTFFT_ACCESS_MODE_READ };

inline std::string get_error_string(Error e) noexcept {
    const char* s = dtfft_get_error_string(static_cast<dtfft_error_t>(e));
    return s ? std::string(s) : std::string("unknown error");
}
inline std::string get_precision_string(Precision p) noexcept {
    return dtfft_get_precision_string(static_cast<dtfft_precision_t>(p));
}
inline std::string get_executor_string(Executor x) noexcept {
    return dtfft_get_executor_string(static_cast<dtfft_executor_t>(x));
}
inline std::string get_backend_string(Backend b) { return dtfft_get_backend_string(static_cast<dtfft_backend_t>(b)); }
inline bool get_backend_pipelined(Backend b) {
    bool pipe = false;
    dtfft_get_backend_pipelined(static_cast<dtfft_backend_t>(b), &pipe);
    return pipe;
}

class Exception final : public std::exception {
public:
    Exception(Error code, std::string msg, const char* file, int line)
        : code_(code), msg_(std::move(msg)), file_(file ? file : ""), line_(line) {
        what_ = "dtFFT Exception: '" + msg_ + "' at " + file_ + ":" + std::to_string(line_);
    }
    const char* what() const noexcept override { return what_.c_str(); }
    Error get_error_code() const noexcept { return code_; }
    const std::string& get_message() const noexcept { return msg_; }
    const std::string& get_file() const noexcept { return file_; }
    int get_line() const noexcept { return line_; }

private:
    Error code_;
    std::string msg_, file_, what_;
    int line_;
};

namespace detail {
inline Error as_error(dtfft_error_t e) noexcept { return static_cast<Error>(e); }
inline void raise_if(Error e, const char* file, int line) {
    if (e != Error::SUCCESS) throw Exception(e, get_error_string(e), file, line);
}
}  // namespace detail
#define DTFFT_CXX_CALL(call) ::dtfft::detail::raise_if((call), __FILE__, __LINE__);

/* Local box of one rank in natural order (x fastest); reference: dtfft.hpp:631-684. */
struct Pencil {
    Pencil() : created_(false) { c_ = dtfft_pencil_t{}; }
    explicit Pencil(dtfft_pencil_t& c_pencil) : created_(true), c_(c_pencil) {}
    explicit Pencil(int32_t n_dims, const int32_t* starts, const int32_t* counts) : created_(true) {
        c_ = dtfft_pencil_t{};
        c_.ndims = static_cast<uint8_t>(n_dims);
        c_.size = 1;
        for (int i = 0; i < n_dims && i < 3; ++i) {
            c_.starts[i] = starts[i];
            c_.counts[i] = counts[i];
            c_.size *= static_cast<size_t>(counts[i]);
        }
    }
    explicit Pencil(const std::vector<int32_t>& starts, const std::vector<int32_t>& counts)
        : Pencil(static_cast<int32_t>(starts.size()), starts.data(), counts.data()) {
        if (starts.size() != counts.size())
            throw Exception(Error::PENCIL_ARRAYS_SIZE_MISMATCH, get_error_string(Error::PENCIL_ARRAYS_SIZE_MISMATCH),
                            __FILE__, __LINE__);
    }
    uint8_t get_ndims() const { return need().ndims; }
    uint8_t get_dim() const { return need().dim; }
    std::vector<int32_t> get_starts() const { return {need().starts, need().starts + need().ndims}; }
    std::vector<int32_t> get_counts() const { return {need().counts, need().counts + need().ndims}; }
    size_t get_size() const { return need().size; }
    const dtfft_pencil_t& c_struct() const { return need(); }

private:
    const dtfft_pencil_t& need() const {
        if (!created_)
            throw Exception(Error::PENCIL_NOT_INITIALIZED, get_error_string(Error::PENCIL_NOT_INITIALIZED), __FILE__,
                            __LINE__);
        return c_;
    }
    bool created_;
    dtfft_pencil_t c_;
};

/* dtfft_config_t with chained setters (dtfft.hpp:710-1070); applied by set_config(). */
struct Config {
    explicit Config() { DTFFT_CXX_CALL(detail::as_error(dtfft_create_config(&config))) }
#define DTFFTB_SETTER(name, type, expr)         \
    Config& set_##name(type v) noexcept {       \
        config.name = (expr);                   \
        return *this;                           \
    }
    DTFFTB_SETTER(enable_log, bool, v)
    DTFFTB_SETTER(enable_z_slab, bool, v)
    DTFFTB_SETTER(enable_y_slab, bool, v)
    DTFFTB_SETTER(n_measure_warmup_iters, int32_t, v)
    DTFFTB_SETTER(n_measure_iters, int32_t, v)
    DTFFTB_SETTER(platform, Platform, static_cast<dtfft_platform_t>(v))
    DTFFTB_SETTER(stream, dtfft_stream_t, v)
    DTFFTB_SETTER(backend, Backend, static_cast<dtfft_backend_t>(v))
    DTFFTB_SETTER(reshape_backend, Backend, static_cast<dtfft_backend_t>(v))
    DTFFTB_SETTER(enable_datatype_backend, bool, v)
    DTFFTB_SETTER(enable_mpi_backends, bool, v)
    DTFFTB_SETTER(enable_pipelined_backends, bool, v)
    DTFFTB_SETTER(enable_rma_backends, bool, v)
    DTFFTB_SETTER(enable_fused_backends, bool, v)
    DTFFTB_SETTER(enable_nccl_backends, bool, v)
    DTFFTB_SETTER(enable_nvshmem_backends, bool, v)
    DTFFTB_SETTER(enable_kernel_autotune, bool, v)
    DTFFTB_SETTER(enable_fourier_reshape, bool, v)
    DTFFTB_SETTER(transpose_mode, TransposeMode, static_cast<dtfft_transpose_mode_t>(v))
    DTFFTB_SETTER(access_mode, AccessMode, static_cast<dtfft_access_mode_t>(v))
#undef DTFFTB_SETTER
    dtfft_config_t c_struct() const { return config; }

protected:
    dtfft_config_t config;
};

inline Error set_config(const Config& config) noexcept {
    const dtfft_config_t c = config.c_struct();
    return detail::as_error(dtfft_set_config(&c));
}

/* Abstract plan (dtfft.hpp:1085-1956).  Non-copyable, movable; destroys the C plan when it dies. */
class Plan {
public:
    Plan(const Plan&) = delete;
    Plan& operator=(const Plan&) = delete;
    Plan(Plan&& o) noexcept : _plan(o._plan) { o._plan = nullptr; }
    virtual ~Plan() noexcept = 0;

    /* value getters: noexcept form + throwing form */
#define DTFFTB_GETTER(name, ctype, cxxtype)                                           \
    Error get_##name(cxxtype* out) const noexcept {                                   \
        ctype v{};                                                                    \
        const Error e = detail::as_error(dtfft_get_##name(_plan, &v));                \
        if (e == Error::SUCCESS && out) *out = static_cast<cxxtype>(v);               \
        return e;                                                                     \
    }                                                                                 \
    cxxtype get_##name() const {                                                      \
        cxxtype v{};                                                                  \
        DTFFT_CXX_CALL(get_##name(&v))                                                \
        return v;                                                                     \
    }
    DTFFTB_GETTER(z_slab_enabled, bool, bool)
    DTFFTB_GETTER(y_slab_enabled, bool, bool)
    DTFFTB_GETTER(alloc_size, size_t, size_t)
    DTFFTB_GETTER(alloc_bytes, size_t, size_t)
    DTFFTB_GETTER(element_size, size_t, size_t)
    DTFFTB_GETTER(aux_size, size_t, size_t)
    DTFFTB_GETTER(aux_bytes, size_t, size_t)
    DTFFTB_GETTER(aux_size_reshape, size_t, size_t)
    DTFFTB_GETTER(aux_bytes_reshape, size_t, size_t)
    DTFFTB_GETTER(aux_size_transpose, size_t, size_t)
    DTFFTB_GETTER(aux_bytes_transpose, size_t, size_t)
    DTFFTB_GETTER(executor, dtfft_executor_t, Executor)
    DTFFTB_GETTER(precision, dtfft_precision_t, Precision)
    DTFFTB_GETTER(stream, dtfft_stream_t, dtfft_stream_t)
#undef DTFFTB_GETTER
    Error get_backend(Backend& b) const noexcept {
        dtfft_backend_t v{};
        const Error e = detail::as_error(dtfft_get_backend(_plan, &v));
        if (e == Error::SUCCESS) b = static_cast<Backend>(v);
        return e;
    }
    Backend get_backend() const {
        Backend b{};
        DTFFT_CXX_CALL(get_backend(b))
        return b;
    }
    Error get_reshape_backend(Backend& b) const noexcept {
        dtfft_backend_t v{};
        const Error e = detail::as_error(dtfft_get_reshape_backend(_plan, &v));
        if (e == Error::SUCCESS) b = static_cast<Backend>(v);
        return e;
    }
    Backend get_reshape_backend() const {
        Backend b{};
        DTFFT_CXX_CALL(get_reshape_backend(b))
        return b;
    }
    Error get_platform(Platform& p) const noexcept {
        dtfft_platform_t v{};
        const Error e = detail::as_error(dtfft_get_platform(_plan, &v));
        if (e == Error::SUCCESS) p = static_cast<Platform>(v);
        return e;
    }
    Platform get_platform() const {
        Platform p{};
        DTFFT_CXX_CALL(get_platform(p))
        return p;
    }

    Error report() const noexcept { return detail::as_error(dtfft_report(_plan)); }

    Error get_pencil(Layout layout, Pencil& pencil) const noexcept {
        dtfft_pencil_t c{};
        const Error e = detail::as_error(dtfft_get_pencil(_plan, static_cast<dtfft_layout_t>(layout), &c));
        if (e == Error::SUCCESS) pencil = Pencil(c);
        return e;
    }
    Pencil get_pencil(Layout layout) const {
        Pencil p;
        DTFFT_CXX_CALL(get_pencil(layout, p))
        return p;
    }

    /* ---- execution ------------------------------------------------------------------------ */
    Error execute(void* in, void* out, Execute type, void* aux = nullptr) const noexcept {
        return detail::as_error(dtfft_execute(_plan, in, out, static_cast<dtfft_execute_t>(type), aux));
    }
    template <typename Tr>
    Tr* execute(void* inout, Execute type, void* aux = nullptr) const {
        DTFFT_CXX_CALL(execute(inout, inout, type, aux))
        return static_cast<Tr*>(inout);
    }
    template <typename T, typename Tr = T>
    Tr* execute(T* inout, Execute type, void* aux = nullptr) const {
        return execute<Tr>(static_cast<void*>(inout), type, aux);
    }
    Error forward(void* in, void* out, void* aux) const noexcept { return execute(in, out, Execute::FORWARD, aux); }
    template <typename Tr>
    Tr* forward(void* inout, void* aux = nullptr) const {
        return execute<Tr>(inout, Execute::FORWARD, aux);
    }
    template <typename T, typename Tr = T>
    Tr* forward(T* inout, void* aux = nullptr) const {
        return execute<T, Tr>(inout, Execute::FORWARD, aux);
    }
    Error backward(void* in, void* out, void* aux) const noexcept { return execute(in, out, Execute::BACKWARD, aux); }
    template <typename Tr>
    Tr* backward(void* inout, void* aux = nullptr) const {
        return execute<Tr>(inout, Execute::BACKWARD, aux);
    }
    template <typename T, typename Tr = T>
    Tr* backward(T* inout, void* aux = nullptr) const {
        return execute<T, Tr>(inout, Execute::BACKWARD, aux);
    }

    Error transpose(void* in, void* out, Transpose type, void* aux = nullptr) const noexcept {
        return detail::as_error(dtfft_transpose(_plan, in, out, static_cast<dtfft_transpose_t>(type), aux));
    }
    Error transpose_start(void* in, void* out, Transpose type, void* aux, dtfft_request_t* request) const noexcept {
        return detail::as_error(
            dtfft_transpose_start(_plan, in, out, static_cast<dtfft_transpose_t>(type), aux, request));
    }
    Error transpose_start(void* in, void* out, Transpose type, dtfft_request_t* request) const noexcept {
        return transpose_start(in, out, type, nullptr, request);
    }
    dtfft_request_t transpose_start(void* in, void* out, Transpose type, void* aux = nullptr) const {
        dtfft_request_t r = nullptr;
        DTFFT_CXX_CALL(transpose_start(in, out, type, aux, &r))
        return r;
    }
    Error transpose_end(dtfft_request_t request) const noexcept {
        return detail::as_error(dtfft_transpose_end(_plan, request));
    }
    Error reshape(void* in, void* out, Reshape type, void* aux = nullptr) const noexcept {
        return detail::as_error(dtfft_reshape(_plan, in, out, static_cast<dtfft_reshape_t>(type), aux));
    }
    Error reshape_start(void* in, void* out, Reshape type, void* aux, dtfft_request_t* request) const noexcept {
        return detail::as_error(dtfft_reshape_start(_plan, in, out, static_cast<dtfft_reshape_t>(type), aux, request));
    }
    Error reshape_start(void* in, void* out, Reshape type, dtfft_request_t* request) const noexcept {
        return reshape_start(in, out, type, nullptr, request);
    }
    dtfft_request_t reshape_start(void* in, void* out, Reshape type, void* aux = nullptr) const {
        dtfft_request_t r = nullptr;
        DTFFT_CXX_CALL(reshape_start(in, out, type, aux, &r))
        return r;
    }
    Error reshape_end(dtfft_request_t request) const noexcept {
        return detail::as_error(dtfft_reshape_end(_plan, request));
    }

    /* ---- sizes ---------------------------------------------------------------------------- */
    Error get_local_sizes(int32_t* in_starts = nullptr, int32_t* in_counts = nullptr, int32_t* out_starts = nullptr,
                          int32_t* out_counts = nullptr, size_t* alloc_size = nullptr) const noexcept {
        return detail::as_error(dtfft_get_local_sizes(_plan, in_starts, in_counts, out_starts, out_counts, alloc_size));
    }
    /* The vectors must already hold ndims entries (reference: dtfft.hpp:1653-1659). */
    Error get_local_sizes(std::vector<int32_t>& in_starts, std::vector<int32_t>& in_counts,
                          std::vector<int32_t>& out_starts, std::vector<int32_t>& out_counts,
                          size_t* alloc_size) const noexcept {
        return get_local_sizes(in_starts.data(), in_counts.data(), out_starts.data(), out_counts.data(), alloc_size);
    }
    Error get_dims(int8_t* ndims, const int32_t* dims[]) const noexcept {
        return detail::as_error(dtfft_get_dims(_plan, ndims, dims));
    }
    std::vector<int32_t> get_dims() const {
        int8_t n = 0;
        const int32_t* p = nullptr;
        DTFFT_CXX_CALL(get_dims(&n, &p))
        return std::vector<int32_t>(p, p + n);
    }
    Error get_grid_dims(int8_t* ndims, const int32_t* grid_dims[]) const noexcept {
        return detail::as_error(dtfft_get_grid_dims(_plan, ndims, grid_dims));
    }
    std::vector<int32_t> get_grid_dims() const {
        int8_t n = 0;
        const int32_t* p = nullptr;
        DTFFT_CXX_CALL(get_grid_dims(&n, &p))
        return std::vector<int32_t>(p, p + n);
    }

    /* ---- memory --------------------------------------------------------------------------- */
    Error mem_alloc(size_t alloc_bytes, void** ptr) const noexcept {
        return detail::as_error(dtfft_mem_alloc(_plan, alloc_bytes, ptr));
    }
    void* mem_alloc(size_t alloc_bytes) const {
        void* p = nullptr;
        DTFFT_CXX_CALL(mem_alloc(alloc_bytes, &p))
        return p;
    }
    template <typename T>
    T* mem_alloc(size_t alloc_bytes) const {
        return static_cast<T*>(mem_alloc(alloc_bytes));
    }
    Error mem_free(void* ptr) const noexcept { return detail::as_error(dtfft_mem_free(_plan, ptr)); }

    Error destroy() noexcept {
        if (!_plan) return Error::SUCCESS;
        return detail::as_error(dtfft_destroy(&_plan));
    }
    dtfft_plan_t c_struct() const { return _plan; }

    /* ---- extensions of dtfft_b200 (no counterpart in the reference) ------------------------ */
    /* NVLINK_FUSED: make a user-allocated device buffer reachable by the peers (collective). */
    Error register_buffer(void* ptr, size_t bytes) const noexcept {
        return detail::as_error(dtfftb_plan_register_buffer(_plan, ptr, bytes));
    }
    Error unregister_buffer(void* ptr) const noexcept {
        return detail::as_error(dtfftb_plan_unregister_buffer(_plan, ptr));
    }
    Error set_overlap(int nchunks, int exchange_ctas = 0) const noexcept {
        return detail::as_error(dtfftb_plan_set_overlap(_plan, nchunks, exchange_ctas));
    }
    Error set_graphs(bool enable) const noexcept { return detail::as_error(dtfftb_plan_set_graphs(_plan, enable ? 1 : 0)); }
    struct Stats {
        int64_t kernel_launches = 0, local_bytes = 0, remote_bytes = 0;
    };
    Stats get_stats() const {
        Stats s;
        DTFFT_CXX_CALL(detail::as_error(dtfftb_plan_get_stats(_plan, &s.kernel_launches, &s.local_bytes, &s.remote_bytes)))
        return s;
    }
    /* NVLINK_FUSED: calls that ran on the NCCL stand-in because a buffer could not be shared through cudaIpc. */
    int64_t get_fallbacks() const {
        int64_t n = 0;
        DTFFT_CXX_CALL(detail::as_error(dtfftb_plan_get_fallbacks(_plan, &n)))
        return n;
    }
    /* How one transposition moves its data on this rank: 0 local kernel, 1 NCCL, 2 direct-store kernel, 3 copy engines,
     * 4 direct-store kernel alone / copy engines in pair pipelines; copies = strided peer copies per execute. */
    struct ExchangeForm {
        int form = 0, copies = 0;
    };
    ExchangeForm get_exchange_form(Transpose transpose_type) const {
        ExchangeForm f;
        DTFFT_CXX_CALL(detail::as_error(dtfftb_plan_get_exchange_form(_plan, static_cast<int>(transpose_type), &f.form, &f.copies)))
        return f;
    }
    /* Non-zero once a device barrier between GPUs has timed out (the plan is dead: every call returns
     * DTFFTB_ERROR_PEER_TIMEOUT). */
    int peer_error() const noexcept { return dtfftb_plan_peer_error(_plan); }

protected:
    Plan() : _plan(nullptr) {}
    dtfft_plan_t _plan;
};
inline Plan::~Plan() noexcept { destroy(); }

class PlanC2C final : public Plan {
public:
    explicit PlanC2C(const std::vector<int32_t>& dims, dtfft_comm_t comm = nullptr,
                     Precision precision = Precision::DOUBLE, Effort effort = Effort::ESTIMATE,
                     Executor executor = Executor::NONE)
        : PlanC2C(static_cast<int8_t>(dims.size()), dims.data(), comm, precision, effort, executor) {}
    explicit PlanC2C(const std::vector<int32_t>& dims, Precision precision, Effort effort = Effort::ESTIMATE,
                     Executor executor = Executor::NONE)
        : PlanC2C(static_cast<int8_t>(dims.size()), dims.data(), nullptr, precision, effort, executor) {}
    explicit PlanC2C(int8_t ndims, const int32_t* dims, dtfft_comm_t comm = nullptr,
                     Precision precision = Precision::DOUBLE, Effort effort = Effort::ESTIMATE,
                     Executor executor = Executor::NONE) {
        DTFFT_CXX_CALL(detail::as_error(dtfft_create_plan_c2c(ndims, dims, comm, static_cast<dtfft_precision_t>(precision),
                                                              static_cast<dtfft_effort_t>(effort),
                                                              static_cast<dtfft_executor_t>(executor), &_plan)))
    }
    explicit PlanC2C(const Pencil& pencil, Precision precision, Effort effort = Effort::ESTIMATE,
                     Executor executor = Executor::NONE)
        : PlanC2C(pencil, nullptr, precision, effort, executor) {}
    explicit PlanC2C(const Pencil& pencil, dtfft_comm_t comm = nullptr, Precision precision = Precision::DOUBLE,
                     Effort effort = Effort::ESTIMATE, Executor executor = Executor::NONE) {
        DTFFT_CXX_CALL(detail::as_error(dtfft_create_plan_c2c_pencil(&pencil.c_struct(), comm,
                                                                     static_cast<dtfft_precision_t>(precision),
                                                                     static_cast<dtfft_effort_t>(effort),
                                                                     static_cast<dtfft_executor_t>(executor), &_plan)))
    }
};

class PlanR2C final : public Plan {
public:
    /* Same argument order as the reference; `executor` must not stay Executor::NONE
     * (Error::R2C_TRANSPOSE_PLAN), exactly like upstream (dtfft.hpp:2041-2126). */
    explicit PlanR2C(const std::vector<int32_t>& dims, dtfft_comm_t comm = nullptr,
                     Precision precision = Precision::DOUBLE, Effort effort = Effort::ESTIMATE,
                     Executor executor = Executor::NONE)
        : PlanR2C(static_cast<int8_t>(dims.size()), dims.data(), comm, precision, effort, executor) {}
    explicit PlanR2C(const std::vector<int32_t>& dims, Precision precision, Effort effort = Effort::ESTIMATE)
        : PlanR2C(static_cast<int8_t>(dims.size()), dims.data(), nullptr, precision, effort, Executor::NONE) {}
    explicit PlanR2C(int8_t ndims, const int32_t* dims, dtfft_comm_t comm = nullptr,
                     Precision precision = Precision::DOUBLE, Effort effort = Effort::ESTIMATE,
                     Executor executor = Executor::NONE) {
        DTFFT_CXX_CALL(detail::as_error(dtfft_create_plan_r2c(ndims, dims, comm, static_cast<dtfft_precision_t>(precision),
                                                              static_cast<dtfft_effort_t>(effort),
                                                              static_cast<dtfft_executor_t>(executor), &_plan)))
    }
    explicit PlanR2C(const Pencil& pencil, Precision precision, Effort effort = Effort::ESTIMATE)
        : PlanR2C(pencil, nullptr, precision, effort, Executor::NONE) {}
    explicit PlanR2C(const Pencil& pencil, dtfft_comm_t comm = nullptr, Precision precision = Precision::DOUBLE,
                     Effort effort = Effort::ESTIMATE, Executor executor = Executor::NONE) {
        DTFFT_CXX_CALL(detail::as_error(dtfft_create_plan_r2c_pencil(&pencil.c_struct(), comm,
                                                                     static_cast<dtfft_precision_t>(precision),
                                                                     static_cast<dtfft_effort_t>(effort),
                                                                     static_cast<dtfft_executor_t>(executor), &_plan)))
    }
};

class PlanR2R final : public Plan {
public:
    explicit PlanR2R(const std::vector<int32_t>& dims, const std::vector<R2RKind>& kinds = std::vector<R2RKind>(),
                     dtfft_comm_t comm = nullptr, Precision precision = Precision::DOUBLE,
                     Effort effort = Effort::ESTIMATE, Executor executor = Executor::NONE)
        : PlanR2R(static_cast<int8_t>(dims.size()), dims.data(), kinds.empty() ? nullptr : kinds.data(), comm, precision,
                  effort, executor) {}
    explicit PlanR2R(const std::vector<int32_t>& dims, Precision precision, Effort effort = Effort::ESTIMATE)
        : PlanR2R(static_cast<int8_t>(dims.size()), dims.data(), nullptr, nullptr, precision, effort, Executor::NONE) {}
    explicit PlanR2R(int8_t ndims, const int32_t* dims, const R2RKind* kinds = nullptr, dtfft_comm_t comm = nullptr,
                     Precision precision = Precision::DOUBLE, Effort effort = Effort::ESTIMATE,
                     Executor executor = Executor::NONE) {
        static_assert(sizeof(R2RKind) == sizeof(dtfft_r2r_kind_t), "R2RKind must alias dtfft_r2r_kind_t");
        DTFFT_CXX_CALL(detail::as_error(dtfft_create_plan_r2r(ndims, dims, reinterpret_cast<const dtfft_r2r_kind_t*>(kinds),
                                                              comm, static_cast<dtfft_precision_t>(precision),
                                                              static_cast<dtfft_effort_t>(effort),
                                                              static_cast<dtfft_executor_t>(executor), &_plan)))
    }
    explicit PlanR2R(const Pencil& pencil, Precision precision, Effort effort = Effort::ESTIMATE)
        : PlanR2R(pencil, nullptr, nullptr, precision, effort, Executor::NONE) {}
    explicit PlanR2R(const Pencil& pencil, const std::vector<R2RKind>& kinds, dtfft_comm_t comm = nullptr,
                     Precision precision = Precision::DOUBLE, Effort effort = Effort::ESTIMATE,
                     Executor executor = Executor::NONE)
        : PlanR2R(pencil, kinds.empty() ? nullptr : kinds.data(), comm, precision, effort, executor) {}
    explicit PlanR2R(const Pencil& pencil, const R2RKind* kinds = nullptr, dtfft_comm_t comm = nullptr,
                     Precision precision = Precision::DOUBLE, Effort effort = Effort::ESTIMATE,
                     Executor executor = Executor::NONE) {
        DTFFT_CXX_CALL(detail::as_error(
            dtfft_create_plan_r2r_pencil(&pencil.c_struct(), reinterpret_cast<const dtfft_r2r_kind_t*>(kinds), comm,
                                         static_cast<dtfft_precision_t>(precision), static_cast<dtfft_effort_t>(effort),
                                         static_cast<dtfft_executor_t>(executor), &_plan)))
    }
};

}  // namespace dtfft
#endif /* DTFFT_B200_HPP */
