// Global configuration: dtfft_config_t + DTFFT_* environment overrides
// (src/dtfft_config.F90:388-483 reads the environment once; :788-802 environment wins over
// the struct passed to dtfft_set_config).
#include <strings.h>

#include <cstdlib>
#include <cstring>

#include "plan.h"

namespace dtfftb {

Config& global_config() {
    static Config c;
    return c;
}

namespace {
bool env_bool(const char* name, bool* out) {
    const char* e = getenv(name);
    if (!e || !*e) return false;
    if (!strcmp(e, "1") || !strcasecmp(e, "true") || !strcasecmp(e, "on") || !strcasecmp(e, "yes")) {
        *out = true;
        return true;
    }
    if (!strcmp(e, "0") || !strcasecmp(e, "false") || !strcasecmp(e, "off") || !strcasecmp(e, "no")) {
        *out = false;
        return true;
    }
    return false;
}
bool env_int(const char* name, int32_t* out) {
    const char* e = getenv(name);
    if (!e || !*e) return false;
    char* end = nullptr;
    long v = strtol(e, &end, 10);
    if (end == e) return false;
    *out = (int32_t)v;
    return true;
}
int backend_from_name(const char* e) {
    if (!e) return 0;
    if (!strcasecmp(e, "nccl")) return BACKEND_NCCL;
    if (!strcasecmp(e, "nccl_pipe")) return BACKEND_NCCL_PIPELINED;
    if (!strcasecmp(e, "nvlink_fused") || !strcasecmp(e, "nvlink") || !strcasecmp(e, "p2p_fused"))
        return BACKEND_NVLINK_FUSED;
    return 0;
}
}  // namespace

Config effective_config() {
    Config c = global_config();
    env_bool("DTFFT_ENABLE_LOG", &c.enable_log);
    env_bool("DTFFT_ENABLE_Z_SLAB", &c.enable_z_slab);
    env_bool("DTFFT_ENABLE_Y_SLAB", &c.enable_y_slab);
    env_int("DTFFT_MEASURE_WARMUP_ITERS", &c.n_measure_warmup_iters);
    env_int("DTFFT_MEASURE_ITERS", &c.n_measure_iters);
    if (int b = backend_from_name(getenv("DTFFT_BACKEND"))) c.backend = b;
    if (int b = backend_from_name(getenv("DTFFT_RESHAPE_BACKEND"))) c.reshape_backend = b;
    env_bool("DTFFT_ENABLE_PIPE", &c.enable_pipelined_backends);
    env_bool("DTFFT_ENABLE_FUSED", &c.enable_fused_backends);
    env_bool("DTFFT_ENABLE_NCCL", &c.enable_nccl_backends);
    env_bool("DTFFT_ENABLE_KERNEL_AUTOTUNE", &c.enable_kernel_autotune);
    env_bool("DTFFT_ENABLE_FOURIER_RESHAPE", &c.enable_fourier_reshape);
    if (const char* e = getenv("DTFFT_PLATFORM")) {
        if (!strcasecmp(e, "cuda")) c.platform = 2;
        if (!strcasecmp(e, "host")) c.platform = 1;
    }
    return c;
}

}  // namespace dtfftb
