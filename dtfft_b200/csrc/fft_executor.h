// cuFFT executor of the plan layer: the B200 stand-in for the reference's cufft_executor
// (src/interfaces/fft/cufft/dtfft_executor_cufft_m.F90:52-125) behind abstract_executor's contract
// (src/dtfft_abstract_executor.F90:67-216), plus batch-range execution for the stage overlap.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <cufftXt.h>

#include <map>

#include "geometry.h"

namespace dtfftb {

class FftExecutor {  // cufft_executor, src/interfaces/fft/cufft/dtfft_executor_cufft_m.F90:52-125
public:
    ~FftExecutor() { destroy(); }
    // fft_rank 1 or 2 along the fastest axis (axes) of `cpx` (and `real` for R2C)
    int create(int fft_rank, bool r2c, int precision, const Pencil* real, const Pencil& cpx, cudaStream_t stream);
    // The deferred create_private of abstract_executor, argument for argument
    // (src/dtfft_abstract_executor.F90:67-84): transform sizes slowest first like cuFFT, unit
    // stride, `idist` / `odist` elements between consecutive transforms.
    int create_raw(int fft_rank, bool r2c, int precision, long long idist, long long odist, long long how_many,
                   const int* fft_sizes, const int* inembed, const int* onembed, cudaStream_t stream);
    int execute(void* a, void* b, int sign);  // sign -1 forward, +1 backward
    // Stage overlap (no reference counterpart): a contiguous range of the batch,
    // [first, first + count) of how_many() transforms.  prepare_range builds the cuFFT plan of a
    // given count ahead of time (plans are cached by count).
    long long how_many() const { return how_many_; }
    int prepare_range(long long count);
    int execute_range(void* a, void* b, int sign, long long first, long long count);
    bool created() const { return created_; }
    void destroy();

private:
    struct Handles {
        cufftHandle fwd = 0, bwd = 0;
    };
    int make_plans(long long how_many, Handles* h);
    Handles whole_;
    std::map<long long, Handles> by_batch_;  // chunk plans keyed by their batch count
    cudaStream_t stream_ = nullptr;
    int rank_ = 1, n_[2] = {1, 1}, inembed_[2] = {1, 1}, onembed_[2] = {1, 1};
    long long idist_ = 1, odist_ = 1, how_many_ = 0;
    size_t in_bytes_ = 16, out_bytes_ = 16;  // element bytes on the forward input / output side
    int precision_ = 1;
    bool created_ = false, r2c_ = false;
};

}  // namespace dtfftb
