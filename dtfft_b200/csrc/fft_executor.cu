// cuFFT executor (see fft_executor.h for the reference citations).
#include "fft_executor.h"

#include "errors.h"

namespace dtfftb {

namespace {
int cufft_error(cufftResult r) { return r == CUFFT_SUCCESS ? 0 : DTFFTB_ERROR_CUDA_BASE - 5000 - (int)r; }
}  // namespace

// ======================================================================================
// FFT executor (cuFFT)
// ======================================================================================
int FftExecutor::make_plans(long long how_many, Handles* h) {
    cufftResult cr;
    const bool sp = precision_ == DTFFT_SINGLE;
    if (!r2c_) {  // dtfft_executor_cufft_m.F90:70-79
        cr = cufftPlanMany(&h->fwd, rank_, n_, inembed_, 1, (int)idist_, onembed_, 1, (int)odist_,
                           sp ? CUFFT_C2C : CUFFT_Z2Z, (int)how_many);
        if (cr != CUFFT_SUCCESS) return cufft_error(cr);
        h->bwd = h->fwd;
    } else {  // :80-93
        cr = cufftPlanMany(&h->fwd, rank_, n_, inembed_, 1, (int)idist_, onembed_, 1, (int)odist_,
                           sp ? CUFFT_R2C : CUFFT_D2Z, (int)how_many);
        if (cr != CUFFT_SUCCESS) return cufft_error(cr);
        cr = cufftPlanMany(&h->bwd, rank_, n_, onembed_, 1, (int)odist_, inembed_, 1, (int)idist_,
                           sp ? CUFFT_C2R : CUFFT_Z2D, (int)how_many);
        if (cr != CUFFT_SUCCESS) return cufft_error(cr);
    }
    cr = cufftSetStream(h->fwd, stream_);
    if (cr != CUFFT_SUCCESS) return cufft_error(cr);
    if (h->bwd != h->fwd) {
        cr = cufftSetStream(h->bwd, stream_);
        if (cr != CUFFT_SUCCESS) return cufft_error(cr);
    }
    return DTFFT_SUCCESS;
}

int FftExecutor::create(int fft_rank, bool r2c, int precision, const Pencil* real, const Pencil& cpx,
                        cudaStream_t stream) {
    // abstract_executor%create, src/dtfft_abstract_executor.F90:115-216
    int n[2] = {1, 1}, inembed[2] = {1, 1}, onembed[2] = {1, 1};
    const Pencil& base = r2c ? *real : cpx;
    if (fft_rank == 1) {
        n[0] = base.counts[0];
        inembed[0] = n[0];
        onembed[0] = cpx.counts[0];
    } else {
        n[0] = base.counts[1], n[1] = base.counts[0];
        inembed[0] = n[0], inembed[1] = n[1];
        onembed[0] = cpx.counts[1], onembed[1] = cpx.counts[0];
    }
    long long idist = 1, odist = 1;
    for (int i = 0; i < fft_rank; ++i) idist *= inembed[i], odist *= onembed[i];
    const long long how_many = (idist == 0 || base.size() == 0) ? 0 : base.size() / idist;
    return create_raw(fft_rank, r2c, precision, idist, odist, how_many, n, inembed, onembed, stream);
}

int FftExecutor::create_raw(int fft_rank, bool r2c, int precision, long long idist, long long odist, long long how_many,
                            const int* fft_sizes, const int* inembed, const int* onembed, cudaStream_t stream) {
    destroy();
    if (fft_rank != 1 && fft_rank != 2) return DTFFTB_ERROR_INTERNAL;
    r2c_ = r2c;
    rank_ = fft_rank;
    precision_ = precision;
    stream_ = stream;
    for (int i = 0; i < fft_rank; ++i) n_[i] = fft_sizes[i], inembed_[i] = inembed[i], onembed_[i] = onembed[i];
    idist_ = idist, odist_ = odist;
    how_many_ = how_many;
    if (idist_ <= 0 || how_many_ <= 0) {  // rank without data: no FFT needed
        how_many_ = 0;
        return DTFFT_SUCCESS;
    }
    const size_t cb = precision == DTFFT_SINGLE ? 8 : 16;
    out_bytes_ = cb;
    in_bytes_ = r2c ? cb / 2 : cb;
    int rc = make_plans(how_many_, &whole_);
    if (rc) return rc;
    created_ = true;
    return DTFFT_SUCCESS;
}

int FftExecutor::execute(void* a, void* b, int sign) {
    if (!created_) return DTFFT_SUCCESS;
    cufftResult cr = cufftXtExec(sign < 0 ? whole_.fwd : whole_.bwd, a, b, sign < 0 ? CUFFT_FORWARD : CUFFT_INVERSE);
    return cufft_error(cr);
}

int FftExecutor::prepare_range(long long count) {
    if (!created_ || count <= 0 || count >= how_many_ || by_batch_.count(count)) return DTFFT_SUCCESS;
    Handles h;
    int rc = make_plans(count, &h);
    if (rc) return rc;
    by_batch_[count] = h;
    return DTFFT_SUCCESS;
}

int FftExecutor::execute_range(void* a, void* b, int sign, long long first, long long count) {
    if (!created_ || count <= 0) return DTFFT_SUCCESS;
    if (first < 0 || first + count > how_many_) return DTFFTB_ERROR_INTERNAL;
    const Handles* h = &whole_;
    if (count != how_many_) {
        int rc = prepare_range(count);
        if (rc) return rc;
        h = &by_batch_.find(count)->second;
    }
    // forward reads the `idist` side and writes the `odist` side; backward the other way round
    const size_t a_off = (size_t)first * (size_t)(sign < 0 ? idist_ : odist_) * (sign < 0 ? in_bytes_ : out_bytes_);
    const size_t b_off = (size_t)first * (size_t)(sign < 0 ? odist_ : idist_) * (sign < 0 ? out_bytes_ : in_bytes_);
    cufftResult cr = cufftXtExec(sign < 0 ? h->fwd : h->bwd, static_cast<char*>(a) + a_off, static_cast<char*>(b) + b_off,
                                 sign < 0 ? CUFFT_FORWARD : CUFFT_INVERSE);
    return cufft_error(cr);
}

void FftExecutor::destroy() {
    auto drop = [](Handles& h) {
        if (h.fwd) cufftDestroy(h.fwd);
        if (h.bwd && h.bwd != h.fwd) cufftDestroy(h.bwd);
        h = Handles{};
    };
    drop(whole_);
    for (auto& kv : by_batch_) drop(kv.second);
    by_batch_.clear();
    created_ = false;
}

}  // namespace dtfftb
