/* C callers of the public plan API on a box WITHOUT a GPU: argument validation, error codes,
 * configuration, and host-metadata-only ("dry") plans.  Plain C11, links only libdtfft_b200.so.
 * Modelled on the reference's C tests (tests/c/test_c2c_3d_c.c:143-179: dims / grid / local
 * sizes expectations; tests/fortran/test_pencils_f.F90:62-222: error codes).
 * Prints "api_host OK" and exits 0 when every check holds. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dtfft_b200_api.h"

static int failures = 0;
#define EXPECT(cond)                                                        \
    do {                                                                    \
        if (!(cond)) {                                                      \
            fprintf(stderr, "%s:%d: check failed: %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                     \
        }                                                                   \
    } while (0)

int main(void) {
    EXPECT(dtfft_get_version() == DTFFT_VERSION_CODE);
    EXPECT(DTFFT_VERSION(3, 2, 0) == 302000);
    EXPECT(dtfft_get_error_string(DTFFT_SUCCESS) != NULL);
    EXPECT(strlen(dtfft_get_error_string(DTFFT_ERROR_INVALID_AUX)) > 0);
    EXPECT(strcmp(dtfft_get_backend_string(DTFFT_BACKEND_NCCL), dtfft_get_backend_string(DTFFT_BACKEND_NVLINK_FUSED)) != 0);
    bool pipe = false;
    EXPECT(dtfft_get_backend_pipelined(DTFFT_BACKEND_NCCL_PIPELINED, &pipe) == DTFFT_SUCCESS && pipe);
    EXPECT(dtfft_get_backend_pipelined(DTFFT_BACKEND_NCCL, &pipe) == DTFFT_SUCCESS && !pipe);

    /* no plan: every entry point answers PLAN_NOT_CREATED */
    size_t n = 0;
    char dummy[8];
    EXPECT(dtfft_execute(NULL, dummy, dummy, DTFFT_EXECUTE_FORWARD, NULL) == DTFFT_ERROR_PLAN_NOT_CREATED);
    EXPECT(dtfft_transpose(NULL, dummy, dummy, DTFFT_TRANSPOSE_X_TO_Y, NULL) == DTFFT_ERROR_PLAN_NOT_CREATED);
    EXPECT(dtfft_get_alloc_size(NULL, &n) == DTFFT_ERROR_PLAN_NOT_CREATED);
    dtfft_plan_t none = NULL;
    EXPECT(dtfft_destroy(&none) != DTFFT_SUCCESS);

    /* argument validation happens before any device is touched (dtfft_plan.F90:1905-1980) */
    dtfft_plan_t plan = NULL;
    int32_t dims3[3] = {64, 48, 40};
    int32_t bad[3] = {64, -1, 40};
    EXPECT(dtfft_create_plan_c2c(3, NULL, NULL, DTFFT_DOUBLE, DTFFT_ESTIMATE, DTFFT_EXECUTOR_NONE, &plan) == DTFFT_ERROR_INVALID_USAGE);
    EXPECT(dtfft_create_plan_c2c(4, dims3, NULL, DTFFT_DOUBLE, DTFFT_ESTIMATE, DTFFT_EXECUTOR_NONE, &plan) == DTFFT_ERROR_INVALID_N_DIMENSIONS);
    EXPECT(dtfft_create_plan_c2c(1, dims3, NULL, DTFFT_DOUBLE, DTFFT_ESTIMATE, DTFFT_EXECUTOR_NONE, &plan) == DTFFT_ERROR_INVALID_N_DIMENSIONS);
    EXPECT(dtfft_create_plan_c2c(3, bad, NULL, DTFFT_DOUBLE, DTFFT_ESTIMATE, DTFFT_EXECUTOR_NONE, &plan) == DTFFT_ERROR_INVALID_DIMENSION_SIZE);
    EXPECT(dtfft_create_plan_c2c(3, dims3, NULL, (dtfft_precision_t)7, DTFFT_ESTIMATE, DTFFT_EXECUTOR_NONE, &plan) == DTFFT_ERROR_INVALID_PRECISION);
    EXPECT(dtfft_create_plan_c2c(3, dims3, NULL, DTFFT_DOUBLE, (dtfft_effort_t)9, DTFFT_EXECUTOR_NONE, &plan) == DTFFT_ERROR_INVALID_EFFORT);
    EXPECT(dtfft_create_plan_c2c(3, dims3, NULL, DTFFT_DOUBLE, DTFFT_ESTIMATE, (dtfft_executor_t)42, &plan) == DTFFT_ERROR_INVALID_EXECUTOR);
    /* FFTW / MKL / VkFFT executors do not exist on this path */
    EXPECT(dtfft_create_plan_c2c(3, dims3, NULL, DTFFT_DOUBLE, DTFFT_ESTIMATE, DTFFT_EXECUTOR_FFTW3, &plan) == DTFFT_ERROR_INVALID_PLATFORM_EXECUTOR);
    EXPECT(dtfft_create_plan_r2c(3, dims3, NULL, DTFFT_DOUBLE, DTFFT_ESTIMATE, DTFFT_EXECUTOR_NONE, &plan) == DTFFT_ERROR_R2C_TRANSPOSE_PLAN);
    dtfft_r2r_kind_t kinds[3] = {DTFFT_DCT_2, DTFFT_DCT_2, DTFFT_DCT_2};
    EXPECT(dtfft_create_plan_r2r(3, dims3, NULL, NULL, DTFFT_DOUBLE, DTFFT_ESTIMATE, DTFFT_EXECUTOR_CUFFT, &plan) == DTFFT_ERROR_MISSING_R2R_KINDS);
    EXPECT(dtfft_create_plan_r2r(3, dims3, kinds, NULL, DTFFT_DOUBLE, DTFFT_ESTIMATE, DTFFT_EXECUTOR_CUFFT, &plan) == DTFFT_ERROR_R2R_FFT_NOT_SUPPORTED);
    EXPECT(dtfft_create_plan_c2c_pencil(NULL, NULL, DTFFT_DOUBLE, DTFFT_ESTIMATE, DTFFT_EXECUTOR_NONE, &plan) == DTFFT_ERROR_INVALID_USAGE);
    dtfft_pencil_t empty;
    memset(&empty, 0, sizeof(empty));
    EXPECT(dtfft_create_plan_c2c_pencil(&empty, NULL, DTFFT_DOUBLE, DTFFT_ESTIMATE, DTFFT_EXECUTOR_NONE, &plan) == DTFFT_ERROR_PENCIL_NOT_INITIALIZED);

    /* a valid request needs a device: success on a GPU box, GPU_NOT_SET elsewhere; never a CPU plan */
    dtfft_error_t rc = dtfft_create_plan_c2c(3, dims3, NULL, DTFFT_DOUBLE, DTFFT_ESTIMATE, DTFFT_EXECUTOR_NONE, &plan);
    EXPECT(rc == DTFFT_SUCCESS || rc == DTFFT_ERROR_GPU_NOT_SET);
    if (rc == DTFFT_SUCCESS) EXPECT(dtfft_destroy(&plan) == DTFFT_SUCCESS && plan == NULL);

    /* configuration: defaults of src/dtfft_config.F90:644-669 (platform is always CUDA here) */
    dtfft_config_t conf;
    EXPECT(dtfft_create_config(&conf) == DTFFT_SUCCESS);
    EXPECT(conf.enable_z_slab && !conf.enable_y_slab && !conf.enable_log);
    EXPECT(conf.n_measure_warmup_iters == 2 && conf.n_measure_iters == 5);
    EXPECT(conf.platform == DTFFT_PLATFORM_CUDA && conf.stream == NULL);
    EXPECT(conf.enable_pipelined_backends && conf.enable_nccl_backends && !conf.enable_kernel_autotune);
    conf.n_measure_iters = 0;
    EXPECT(dtfft_set_config(&conf) == DTFFT_ERROR_INVALID_MEASURE_ITERS);
    conf.n_measure_iters = 5;
    conf.platform = DTFFT_PLATFORM_HOST; /* there is no CPU path */
    EXPECT(dtfft_set_config(&conf) == DTFFT_ERROR_INVALID_PLATFORM);
    conf.platform = DTFFT_PLATFORM_CUDA;
    conf.enable_z_slab = false;
    EXPECT(dtfft_set_config(&conf) == DTFFT_SUCCESS);

    /* dry plan (host metadata only): one rank owns everything */
    EXPECT(dtfftb_plan_create_dry(0, 3, dims3, NULL, NULL, DTFFT_DOUBLE, DTFFT_EXECUTOR_NONE, &plan) == DTFFT_SUCCESS);
    int8_t nd = 0;
    const int32_t* d = NULL;
    EXPECT(dtfft_get_dims(plan, &nd, &d) == DTFFT_SUCCESS && nd == 3 && d[0] == 64 && d[1] == 48 && d[2] == 40);
    const int32_t* g = NULL;
    EXPECT(dtfft_get_grid_dims(plan, NULL, &g) == DTFFT_SUCCESS && g[0] == 1 && g[1] == 1 && g[2] == 1);
    int32_t is[3], ic[3], os[3], oc[3];
    size_t alloc = 0, esize = 0, bytes = 0;
    EXPECT(dtfft_get_local_sizes(plan, is, ic, os, oc, &alloc) == DTFFT_SUCCESS);
    EXPECT(is[0] == 0 && is[1] == 0 && is[2] == 0 && ic[0] == 64 && ic[1] == 48 && ic[2] == 40);
    EXPECT(oc[0] == 40 && oc[1] == 64 && oc[2] == 48); /* Z pencil: (z, x, y) */
    EXPECT(alloc == (size_t)64 * 48 * 40);
    EXPECT(dtfft_get_element_size(plan, &esize) == DTFFT_SUCCESS && esize == 16);
    EXPECT(dtfft_get_alloc_bytes(plan, &bytes) == DTFFT_SUCCESS && bytes == alloc * 16);
    dtfft_pencil_t p;
    EXPECT(dtfft_get_pencil(plan, DTFFT_LAYOUT_Y_PENCILS, &p) == DTFFT_SUCCESS);
    EXPECT(p.dim == 2 && p.ndims == 3 && p.counts[0] == 48 && p.counts[1] == 40 && p.counts[2] == 64 && p.size == alloc);
    EXPECT(dtfft_get_pencil(plan, (dtfft_layout_t)99, &p) == DTFFT_ERROR_INVALID_LAYOUT);
    EXPECT(dtfft_get_pencil(plan, DTFFT_LAYOUT_X_BRICKS, &p) != DTFFT_SUCCESS); /* no reshape in this plan */
    bool z = true;
    EXPECT(dtfft_get_z_slab_enabled(plan, &z) == DTFFT_SUCCESS && !z);
    void* ptr = NULL;
    EXPECT(dtfft_mem_alloc(plan, 1024, &ptr) == DTFFT_ERROR_GPU_NOT_SET); /* a dry plan never touches a device */
    EXPECT(dtfft_execute(plan, dummy, dummy + 1, DTFFT_EXECUTE_FORWARD, NULL) == DTFFT_ERROR_GPU_NOT_SET);
    /* async requests: nothing was started, so nothing can be retired (CHECK_REQUEST, dtfft_plan.F90:75-84) */
    EXPECT(dtfft_transpose_end(plan, NULL) == DTFFT_ERROR_INVALID_REQUEST);
    EXPECT(dtfft_reshape_end(plan, (dtfft_request_t)dummy) == DTFFT_ERROR_INVALID_REQUEST);
    EXPECT(dtfft_transpose_start(plan, dummy, dummy + 1, DTFFT_TRANSPOSE_X_TO_Y, NULL, NULL) == DTFFT_ERROR_INVALID_USAGE);
    EXPECT(dtfft_destroy(&plan) == DTFFT_SUCCESS && plan == NULL);

    /* dry R2C plan: the complex side has nx/2+1 points along x */
    EXPECT(dtfftb_plan_create_dry(1, 3, dims3, NULL, NULL, DTFFT_SINGLE, DTFFT_EXECUTOR_CUFFT, &plan) == DTFFT_SUCCESS);
    EXPECT(dtfft_get_local_sizes(plan, is, ic, os, oc, &alloc) == DTFFT_SUCCESS);
    EXPECT(ic[0] == 64 && oc[0] == 40 && oc[1] == 33 && oc[2] == 48);
    EXPECT(dtfft_get_element_size(plan, &esize) == DTFFT_SUCCESS && esize == 4);
    EXPECT(dtfft_destroy(&plan) == DTFFT_SUCCESS);

    /* restore the defaults for whoever shares the process */
    dtfft_create_config(&conf);
    dtfft_set_config(&conf);
    if (failures) {
        fprintf(stderr, "api_host: %d check(s) failed\n", failures);
        return 1;
    }
    printf("api_host OK\n");
    return 0;
}
