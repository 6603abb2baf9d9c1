// Block descriptors shared by the host-side geometry builder and the sm_100a kernels.
//
// Every dtFFT reshape kernel kind (src/dtfft_abstract_kernel.F90:59-98 of the reference)
// is a strided move of one or more 3-D boxes of opaque 4/8/16-byte elements.  Instead of
// generating one CUDA-C string per kind at run time (reference:
// src/dtfft_nvrtc_module.F90:434-583) we describe each peer's box by strides and run one
// of two ahead-of-time compiled kernel families over a table of boxes:
//   * family T ("transpose"): input contiguous along axis a, output contiguous along
//     axis b (a != b) -> staged through a shared-memory tile;
//   * family R ("rows"): both sides contiguous along axis a -> vectorised row copy.
// One launch covers ALL peers (the reference launches once per peer,
// src/dtfft_kernel_device.F90:167-174).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define DTFFTB_HD __host__ __device__ __forceinline__
#else
#define DTFFTB_HD inline
#endif

namespace dtfftb {

// Division of a 31-bit dividend by a run-time constant as multiply-high + shift
// (Granlund-Montgomery round-up method).  Work-item decoding uses it instead of
// hardware-emulated integer division.
struct FastDiv {
    unsigned mul = 0, shr = 0, div = 1;
    static inline FastDiv make(unsigned d) {
        FastDiv f;
        f.div = d;
        if (d <= 1) return f;  // mul == 0 marks "divide by one"
        unsigned lg = 0;       // ceil(log2 d)
        while ((1ull << lg) < d) ++lg;
        unsigned p = 31 + lg;
        f.mul = (unsigned)(((1ull << p) + d - 1) / d);
        f.shr = p - 32;
        return f;
    }
};

struct BlockDesc {
    long long in_off;     // offset of the box origin in `in`, in kernel units
    long long out_off;    // offset of the box origin in `out` (or in out_base)
    void* out_base;       // nullptr -> use the launch's `out`; else a (peer-mapped) base pointer
    const void* in_base;  // nullptr -> use the launch's `in`
    long long is1, is2;   // input strides of axes b, c   (axis a has stride 1)
    long long os0, os1, os2;  // output strides of axes a, b, c (family R: os0 == 1)
    long long item_begin; // first work item of this block in the flattened item space
    long long shuffle;    // blocks[0] only: > 1 -> CTA i works on item (i * shuffle) mod total (coprime multiplier)
    const unsigned long long* abort;  // blocks[0] only: non-null -> sticky peer-error word; when set the kernel stores nothing
    int n0, n1, n2;       // extents along a, b, c
    int tiles0, tiles1;   // tiles along a and b
    int bshift;           // family T: tiles along b start `bshift` elements BEFORE the box, so that every tile boundary falls
                          // on a 128-byte line of the destination (0 = box origin already aligned, or rows not alignable)
    FastDiv div0, div1;   // fast division by tiles0 / tiles1
};

}  // namespace dtfftb
