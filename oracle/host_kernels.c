/* Oracle (TEST INFRASTRUCTURE, never linked into the product): plain-C restatement of the
 * reference's CPU kernels, src/include/_dtfft_kernel_host_routines.inc ("*_write" loops,
 * OpenMP collapse as under DTFFT_WITH_OPENMP) and of the blocked whole-pencil permutes of
 * src/include/_dtfft_kernel_host_block_routines.inc.  Used (a) as a second checker beside the
 * numpy oracle and (b) as the CPU baseline timed by bench.py ("kind": "port": the reference
 * itself is Fortran+MPI and cannot be built in this image).  0-based indices; elements are
 * opaque 4/8/16-byte words, so every result is bit-exact.
 *
 * Build: make -C oracle   ->  oracle/_build/liboracle_host.so
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { uint64_t a, b; } w16_t;

enum {
    K_PACK = 1, K_COPY_PIPELINED = 2, K_UNPACK = 3, K_COPY = 4, K_UNPACK_PIPELINED = 5, K_PACK_PIPELINED = 6,
    K_PERMUTE_FORWARD = 7, K_PERMUTE_BACKWARD = 8, K_PERMUTE_BACKWARD_START = 9, K_PERMUTE_BACKWARD_END = 10,
    K_PERMUTE_BACKWARD_END_PIPELINED = 11, K_PACK_FORWARD = 12, K_PACK_BACKWARD = 13
};

#define DEFINE_KERNELS(T, SFX)                                                                              \
    /* .inc:140-196  out(y,z,x) <- in(x,y,z) */                                                             \
    static void permute_forward_##SFX(const T* in, T* out, int64_t nx, int64_t ny, int64_t nz) {            \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                            \
        for (int64_t x = 0; x < nx; ++x)                                                                    \
            for (int64_t z = 0; z < nz; ++z) {                                                              \
                const T* ip = in + z * nx * ny + x;                                                         \
                T* op = out + x * ny * nz + z * ny;                                                         \
                for (int64_t y = 0; y < ny; ++y) op[y] = ip[y * nx];                                        \
            }                                                                                               \
    }                                                                                                       \
    /* .inc:262-302  out(z,x,y) <- in(x,y,z) */                                                             \
    static void permute_backward_##SFX(const T* in, T* out, int64_t nx, int64_t ny, int64_t nz) {           \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                            \
        for (int64_t y = 0; y < ny; ++y)                                                                    \
            for (int64_t x = 0; x < nx; ++x) {                                                              \
                const T* ip = in + y * nx + x;                                                              \
                T* op = out + y * nz * nx + x * nz;                                                         \
                for (int64_t z = 0; z < nz; ++z) op[z] = ip[z * nx * ny];                                   \
            }                                                                                               \
    }                                                                                                       \
    /* .inc:351-390  out(z,y,x) <- in(x,y,z) */                                                             \
    static void permute_backward_start_##SFX(const T* in, T* out, int64_t nx, int64_t ny, int64_t nz) {     \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                            \
        for (int64_t x = 0; x < nx; ++x)                                                                    \
            for (int64_t y = 0; y < ny; ++y) {                                                              \
                const T* ip = in + y * nx + x;                                                              \
                T* op = out + x * nz * ny + y * nz;                                                         \
                for (int64_t z = 0; z < nz; ++z) op[z] = ip[z * nx * ny];                                   \
            }                                                                                               \
    }                                                                                                       \
    /* generic row copy used by unpack (.inc:575-636), pack (:657-721), backward_end (:439-482) */          \
    static void rows_##SFX(const T* in, T* out, int64_t n0, int64_t n1, int64_t n2, int64_t is1,            \
                           int64_t is2, int64_t os1, int64_t os2) {                                         \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                            \
        for (int64_t z = 0; z < n2; ++z)                                                                    \
            for (int64_t y = 0; y < n1; ++y) {                                                              \
                const T* ip = in + z * is2 + y * is1;                                                       \
                T* op = out + z * os2 + y * os1;                                                            \
                for (int64_t x = 0; x < n0; ++x) op[x] = ip[x];                                             \
            }                                                                                               \
    }                                                                                                       \
    /* .inc:1026-1086 */                                                                                    \
    static void pack_forward_##SFX(const T* in, T* out, int64_t nx, int64_t ny, int64_t nxx, int64_t nyy,   \
                                   int64_t nzz) {                                                           \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                            \
        for (int64_t x = 0; x < nxx; ++x)                                                                   \
            for (int64_t z = 0; z < nzz; ++z) {                                                             \
                const T* ip = in + z * nx * ny + x;                                                         \
                T* op = out + x * nyy * nzz + z * nyy;                                                      \
                for (int64_t y = 0; y < nyy; ++y) op[y] = ip[y * nx];                                       \
            }                                                                                               \
    }                                                                                                       \
    /* .inc:1155-1198 */                                                                                    \
    static void pack_backward_##SFX(const T* in, T* out, int64_t nx, int64_t ny, int64_t nxx, int64_t nyy,  \
                                    int64_t nzz) {                                                          \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                            \
        for (int64_t y = 0; y < nyy; ++y)                                                                   \
            for (int64_t x = 0; x < nxx; ++x) {                                                             \
                const T* ip = in + y * nx + x;                                                              \
                T* op = out + y * nzz * nxx + x * nzz;                                                      \
                for (int64_t z = 0; z < nzz; ++z) op[z] = ip[z * nx * ny];                                  \
            }                                                                                               \
    }                                                                                                       \
    static int execute_one_##SFX(int kt, int ndims, const int32_t* dims, const T* in, T* out,              \
                                 const int32_t* l) {                                                        \
        const int64_t nx = dims[0], ny = dims[1], nz = ndims == 3 ? dims[2] : 1;                            \
        int64_t nxx = 0, nyy = 0, nzz = 1, din = 0, dout = 0;                                               \
        if (l) { nxx = l[0]; nyy = l[1]; nzz = ndims == 3 ? l[2] : 1; din = l[3]; dout = l[4]; }            \
        switch (kt) {                                                                                       \
            case K_PERMUTE_FORWARD: permute_forward_##SFX(in, out, nx, ny, nz); return 0;                   \
            case K_PERMUTE_BACKWARD:                                                                        \
                if (ndims == 2) permute_forward_##SFX(in, out, nx, ny, 1);                                  \
                else permute_backward_##SFX(in, out, nx, ny, nz);                                           \
                return 0;                                                                                   \
            case K_PERMUTE_BACKWARD_START: permute_backward_start_##SFX(in, out, nx, ny, nz); return 0;     \
            case K_UNPACK_PIPELINED:                                                                        \
                rows_##SFX(in + din, out + dout, nxx, nyy, nzz, nxx, nxx * nyy, nx, nx * ny); return 0;     \
            case K_PACK_PIPELINED:                                                                          \
                rows_##SFX(in + din, out + dout, nxx, nyy, nzz, nx, nx * ny, nxx, nxx * nyy); return 0;     \
            case K_PERMUTE_BACKWARD_END_PIPELINED:                                                          \
                rows_##SFX(in + din, out + dout, nxx, nyy, nzz, nxx * nzz, nxx, nx, nx * ny); return 0;     \
            case K_PACK_FORWARD: pack_forward_##SFX(in + din, out + dout, nx, ny, nxx, nyy, nzz); return 0; \
            case K_PACK_BACKWARD:                                                                           \
                if (ndims == 2) pack_forward_##SFX(in + din, out + dout, nx, ny, nxx, nyy, 1);              \
                else pack_backward_##SFX(in + din, out + dout, nx, ny, nxx, nyy, nzz);                      \
                return 0;                                                                                   \
            case K_COPY_PIPELINED:                                                                          \
                memcpy(out + dout, in + din, sizeof(T) * (size_t)(nxx * nyy * nzz)); return 0;              \
            case K_COPY: memcpy(out, in, sizeof(T) * (size_t)(nx * ny * nz)); return 0;                     \
            default: return -1;                                                                             \
        }                                                                                                   \
    }

DEFINE_KERNELS(uint32_t, 4)
DEFINE_KERNELS(uint64_t, 8)
DEFINE_KERNELS(w16_t, 16)

static int looped(int kt) {
    switch (kt) {
        case K_PACK: return K_PACK_PIPELINED;
        case K_UNPACK: return K_UNPACK_PIPELINED;
        case K_PERMUTE_BACKWARD_END: return K_PERMUTE_BACKWARD_END_PIPELINED;
        default: return 0;
    }
}

/* neighbor_data: 5 x P column-major (row per neighbour in C order); neighbor is 1-based, 0 = all. */
int oracle_kernel_execute(int kernel_type, int ndims, const int32_t* dims, int es, const void* in, void* out,
                          const int32_t* neighbor_data, int n_neighbors, int neighbor) {
    for (int i = 0; i < ndims; ++i)
        if (dims[i] == 0) return 0;
    int per = looped(kernel_type);
    int first = neighbor > 0 ? neighbor - 1 : 0;
    int last = neighbor > 0 ? neighbor : (per ? n_neighbors : 1);
    int kt = per ? per : kernel_type;
    for (int n = first; n < last; ++n) {
        const int32_t* l = neighbor_data ? neighbor_data + 5 * n : 0;
        if (l && (int64_t)l[0] * l[1] * (ndims == 3 ? l[2] : 1) == 0) continue;
        int rc;
        if (es == 4) rc = execute_one_4(kt, ndims, dims, (const uint32_t*)in, (uint32_t*)out, l);
        else if (es == 8) rc = execute_one_8(kt, ndims, dims, (const uint64_t*)in, (uint64_t*)out, l);
        else if (es == 16) rc = execute_one_16(kt, ndims, dims, (const w16_t*)in, (w16_t*)out, l);
        else return -2;
        if (rc) return rc;
    }
    return 0;
}

/* Blocked variants of the three whole-pencil permutes, src/include/_dtfft_kernel_host_block_routines.inc
 * ("*_write" order: the block loops and the loops inside a block run in the order of the OUTPUT axes,
 * :100-143 forward, :204-232 backward, :290-318 backward_start; BLOCK_SIZE in {4, 8, 16, 32, 64},
 * _dtfft_kernel_host_routines.inc:1247-1255).  The reference's host kernel picks among the unblocked
 * loops and these by timing (kernel_host create); bench.py's CPU legs do the same. */
#define DEFINE_BLOCKED(T, SFX)                                                                              \
    static void permute_forward_blk_##SFX(const T* in, T* out, int64_t nx, int64_t ny, int64_t nz,          \
                                          int64_t B) {                                                      \
        _Pragma("omp parallel for collapse(3) schedule(static)")                                            \
        for (int64_t xb = 0; xb < nx; xb += B)                                                              \
            for (int64_t zb = 0; zb < nz; zb += B)                                                          \
                for (int64_t yb = 0; yb < ny; yb += B) {                                                    \
                    const int64_t xe = xb + B < nx ? xb + B : nx, ze = zb + B < nz ? zb + B : nz;           \
                    const int64_t ye = yb + B < ny ? yb + B : ny;                                           \
                    for (int64_t x = xb; x < xe; ++x)                                                       \
                        for (int64_t z = zb; z < ze; ++z)                                                   \
                            for (int64_t y = yb; y < ye; ++y)                                               \
                                out[x * ny * nz + z * ny + y] = in[z * nx * ny + y * nx + x];               \
                }                                                                                           \
    }                                                                                                       \
    static void permute_backward_blk_##SFX(const T* in, T* out, int64_t nx, int64_t ny, int64_t nz,         \
                                           int64_t B) {                                                     \
        _Pragma("omp parallel for collapse(3) schedule(static)")                                            \
        for (int64_t yb = 0; yb < ny; yb += B)                                                              \
            for (int64_t xb = 0; xb < nx; xb += B)                                                          \
                for (int64_t zb = 0; zb < nz; zb += B) {                                                    \
                    const int64_t xe = xb + B < nx ? xb + B : nx, ze = zb + B < nz ? zb + B : nz;           \
                    const int64_t ye = yb + B < ny ? yb + B : ny;                                           \
                    for (int64_t y = yb; y < ye; ++y)                                                       \
                        for (int64_t x = xb; x < xe; ++x)                                                   \
                            for (int64_t z = zb; z < ze; ++z)                                               \
                                out[y * nz * nx + x * nz + z] = in[z * nx * ny + y * nx + x];               \
                }                                                                                           \
    }                                                                                                       \
    static void permute_backward_start_blk_##SFX(const T* in, T* out, int64_t nx, int64_t ny, int64_t nz,   \
                                                 int64_t B) {                                               \
        _Pragma("omp parallel for collapse(3) schedule(static)")                                            \
        for (int64_t xb = 0; xb < nx; xb += B)                                                              \
            for (int64_t yb = 0; yb < ny; yb += B)                                                          \
                for (int64_t zb = 0; zb < nz; zb += B) {                                                    \
                    const int64_t xe = xb + B < nx ? xb + B : nx, ze = zb + B < nz ? zb + B : nz;           \
                    const int64_t ye = yb + B < ny ? yb + B : ny;                                           \
                    for (int64_t x = xb; x < xe; ++x)                                                       \
                        for (int64_t y = yb; y < ye; ++y)                                                   \
                            for (int64_t z = zb; z < ze; ++z)                                               \
                                out[x * nz * ny + y * nz + z] = in[z * nx * ny + y * nx + x];               \
                }                                                                                           \
    }                                                                                                       \
    static int execute_blocked_##SFX(int kt, int ndims, const int32_t* dims, const T* in, T* out,           \
                                     int64_t B) {                                                           \
        const int64_t nx = dims[0], ny = dims[1], nz = ndims == 3 ? dims[2] : 1;                            \
        switch (kt) {                                                                                       \
            case K_PERMUTE_FORWARD: permute_forward_blk_##SFX(in, out, nx, ny, nz, B); return 0;            \
            case K_PERMUTE_BACKWARD:                                                                        \
                if (ndims == 2) permute_forward_blk_##SFX(in, out, nx, ny, 1, B);                           \
                else permute_backward_blk_##SFX(in, out, nx, ny, nz, B);                                    \
                return 0;                                                                                   \
            case K_PERMUTE_BACKWARD_START: permute_backward_start_blk_##SFX(in, out, nx, ny, nz, B);        \
                return 0;                                                                                   \
            default: return -1;                                                                             \
        }                                                                                                   \
    }

DEFINE_BLOCKED(uint32_t, 4)
DEFINE_BLOCKED(uint64_t, 8)
DEFINE_BLOCKED(w16_t, 16)

/* Whole-pencil permutes with the reference's blocked loops; block in {4, 8, 16, 32, 64}. */
int oracle_kernel_execute_blocked(int kernel_type, int ndims, const int32_t* dims, int es, const void* in, void* out,
                                  int block) {
    if (block != 4 && block != 8 && block != 16 && block != 32 && block != 64) return -3;
    for (int i = 0; i < ndims; ++i)
        if (dims[i] == 0) return 0;
    if (es == 4) return execute_blocked_4(kernel_type, ndims, dims, (const uint32_t*)in, (uint32_t*)out, block);
    if (es == 8) return execute_blocked_8(kernel_type, ndims, dims, (const uint64_t*)in, (uint64_t*)out, block);
    if (es == 16) return execute_blocked_16(kernel_type, ndims, dims, (const w16_t*)in, (w16_t*)out, block);
    return -2;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm of bench.py asks for
 * every host core explicitly (n <= 0: all online processors). */
void oracle_set_num_threads(int n) {
    if (n <= 0) n = omp_get_num_procs();
    omp_set_num_threads(n);
}
