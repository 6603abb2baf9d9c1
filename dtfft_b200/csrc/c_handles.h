// Opaque handle types of the C ABI (include/dtfft_b200.h) shared by the *_api.cu shells.
#pragma once
#include "kernel_object.h"

struct dtfftb_kernel_s {
    dtfftb::Kernel k;
};
