// C++ callers of the plan API on a GPU, through include/dtfft_b200.hpp (the mirror of the
// reference's dtfft.hpp).  Follows the flow of the reference's C/C++ tests
// (tests/c/test_c2c_3d_c.c:181-245, tests/c/test_r2c_2d_cxx.cpp): create -> get_local_sizes ->
// mem_alloc -> fill -> execute forward -> check -> execute backward -> compare with the input
// scaled by prod(dims), error threshold 5 log2(N) 2 eps (tests/test_utils.F90:96,107).
// Transposes are checked BIT-EXACTLY against the index map of the datatype path
// (Z pencil element (z, x, y) == X pencil element (x, y, z)).
// Prints "api_gpu OK" and exits 0 when every check holds.
#include <cuda_runtime.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "dtfft_b200.hpp"

static int failures = 0;
#define EXPECT(cond)                                                                      \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            std::fprintf(stderr, "%s:%d: check failed: %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                                   \
        }                                                                                 \
    } while (0)
#define CUDA_OK(call) EXPECT((call) == cudaSuccess)

using namespace dtfft;

static void transpose_only_3d() {
    const std::vector<int32_t> dims = {40, 33, 27};  // x fastest; odd extents on purpose
    Config conf;
    conf.set_enable_z_slab(false);
    EXPECT(set_config(conf) == Error::SUCCESS);
    PlanC2C plan(dims);  // transpose-only, one rank, double precision
    EXPECT(plan.get_executor() == Executor::NONE && plan.get_precision() == Precision::DOUBLE);
    EXPECT(plan.get_platform() == Platform::CUDA && plan.get_backend() == Backend::NONE);
    EXPECT(plan.get_dims() == dims && plan.get_grid_dims() == std::vector<int32_t>({1, 1, 1}));
    EXPECT(!plan.get_z_slab_enabled());
    std::vector<int32_t> is(3), ic(3), os(3), oc(3);
    size_t alloc = 0;
    EXPECT(plan.get_local_sizes(is, ic, os, oc, &alloc) == Error::SUCCESS);
    EXPECT(ic == dims && oc == std::vector<int32_t>({27, 40, 33}) && alloc == (size_t)40 * 33 * 27);
    EXPECT(plan.get_element_size() == 16 && plan.get_alloc_bytes() == alloc * 16);
    const Pencil zp = plan.get_pencil(Layout::Z_PENCILS);
    EXPECT(zp.get_dim() == 3 && zp.get_counts() == oc && zp.get_size() == alloc);
    EXPECT(plan.report() == Error::SUCCESS);

    using cd = std::complex<double>;
    const size_t n = alloc;
    std::vector<cd> h_in(n), h_out(n), h_back(n);
    std::mt19937_64 rng(1234);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    for (auto& v : h_in) v = cd(u(rng), u(rng));
    cd* a = plan.mem_alloc<cd>(plan.get_alloc_bytes());
    cd* b = plan.mem_alloc<cd>(plan.get_alloc_bytes());
    cd* c = plan.mem_alloc<cd>(plan.get_alloc_bytes());
    cudaStream_t stream = static_cast<cudaStream_t>(plan.get_stream());
    CUDA_OK(cudaMemcpy(a, h_in.data(), n * sizeof(cd), cudaMemcpyHostToDevice));
    EXPECT(plan.forward(a, b, nullptr) == Error::SUCCESS);
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaMemcpy(h_out.data(), b, n * sizeof(cd), cudaMemcpyDeviceToHost));
    const int nx = dims[0], ny = dims[1], nz = dims[2];
    size_t wrong = 0;
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x)
            for (int z = 0; z < nz; ++z) {
                const cd& got = h_out[(size_t)z + (size_t)nz * (x + (size_t)nx * y)];
                const cd& want = h_in[(size_t)x + (size_t)nx * (y + (size_t)ny * z)];
                wrong += std::memcmp(&got, &want, sizeof(cd)) != 0;
            }
    EXPECT(wrong == 0);
    EXPECT(plan.backward(b, c, nullptr) == Error::SUCCESS);
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaMemcpy(h_back.data(), c, n * sizeof(cd), cudaMemcpyDeviceToHost));
    EXPECT(std::memcmp(h_back.data(), h_in.data(), n * sizeof(cd)) == 0);

    // single transpositions and the error conventions of transpose_private (dtfft_plan.F90:695-747)
    CUDA_OK(cudaMemcpy(a, h_in.data(), n * sizeof(cd), cudaMemcpyHostToDevice));
    EXPECT(plan.transpose(a, b, Transpose::X_TO_Y) == Error::SUCCESS);
    EXPECT(plan.transpose(b, a, Transpose::Y_TO_X) == Error::SUCCESS);
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaMemcpy(h_back.data(), a, n * sizeof(cd), cudaMemcpyDeviceToHost));
    EXPECT(std::memcmp(h_back.data(), h_in.data(), n * sizeof(cd)) == 0);
    EXPECT(plan.transpose(a, a, Transpose::X_TO_Y) == Error::INPLACE_TRANSPOSE);
    EXPECT(plan.transpose(a, b, Transpose::X_TO_Z) == Error::INVALID_TRANSPOSE_TYPE);  // no Z-slab in this plan
    EXPECT(plan.transpose(a, b, static_cast<Transpose>(7)) == Error::INVALID_TRANSPOSE_TYPE);
    EXPECT(plan.transpose(a, b, Transpose::X_TO_Y, a) == Error::INVALID_AUX);
    EXPECT(plan.transpose(h_in.data(), b, Transpose::X_TO_Y) == Error::NOT_DEVICE_PTR);
    EXPECT(plan.execute(a, b, static_cast<Execute>(3)) == Error::INVALID_EXECUTE_TYPE);
    EXPECT(plan.reshape(a, b, Reshape::X_BRICKS_TO_PENCILS) == Error::RESHAPE_NOT_SUPPORTED);
    // async requests (dtfft_plan.F90:599-693; CHECK_REQUEST :75-84): a request is retired once, by the
    // plan and the kind of call that started it
    dtfft_request_t req = plan.transpose_start(a, b, Transpose::X_TO_Y);
    EXPECT(req != nullptr);
    EXPECT(plan.reshape_end(req) == Error::INVALID_REQUEST);      // a transposition, not a reshape
    EXPECT(plan.transpose_end(req) == Error::SUCCESS);
    EXPECT(plan.transpose_end(req) == Error::INVALID_REQUEST);    // already retired
    EXPECT(plan.transpose_end(nullptr) == Error::INVALID_REQUEST);
    dtfft_request_t none = nullptr;
    EXPECT(plan.transpose_start(a, a, Transpose::X_TO_Y, &none) == Error::INPLACE_TRANSPOSE && none == nullptr);
    const Plan::Stats st = plan.get_stats();
    EXPECT(st.kernel_launches >= 1 && st.local_bytes == (int64_t)(n * sizeof(cd)) && st.remote_bytes == 0);
    CUDA_OK(cudaStreamSynchronize(stream));
    EXPECT(plan.mem_free(a) == Error::SUCCESS && plan.mem_free(b) == Error::SUCCESS && plan.mem_free(c) == Error::SUCCESS);
    EXPECT(plan.mem_free(c) == Error::FREE_FAILED);
    void* huge = nullptr;  // more than the device has: refused up front (reshape_plan_base.F90:429-432)
    EXPECT(plan.mem_alloc((size_t)1 << 50, &huge) == Error::ALLOC_FAILED && huge == nullptr);
    EXPECT(plan.mem_alloc(0, &huge) == Error::INVALID_ALLOC_BYTES);
    EXPECT(plan.destroy() == Error::SUCCESS && plan.c_struct() == nullptr);
    EXPECT(plan.get_alloc_size(nullptr) == Error::PLAN_NOT_CREATED);
}

static void r2c_2d_float_cufft() {
    const std::vector<int32_t> dims = {66, 40};
    EXPECT(set_config(Config()) == Error::SUCCESS);
    bool thrown = false;
    try {
        PlanR2C bad(dims, Precision::SINGLE);  // R2C needs an executor
    } catch (const Exception& e) {
        thrown = e.get_error_code() == Error::R2C_TRANSPOSE_PLAN && std::strlen(e.what()) > 0;
    }
    EXPECT(thrown);
    PlanR2C plan(dims, nullptr, Precision::SINGLE, Effort::ESTIMATE, Executor::CUFFT);
    std::vector<int32_t> is(2), ic(2), os(2), oc(2);
    size_t alloc = 0;
    EXPECT(plan.get_local_sizes(is, ic, os, oc, &alloc) == Error::SUCCESS);
    EXPECT(ic == dims && oc == std::vector<int32_t>({40, 34}));  // Y pencil of the complex side: (y, x/2+1)
    EXPECT(plan.get_element_size() == 4);
    const size_t n_real = (size_t)66 * 40;
    std::vector<float> h_in(n_real), h_back(n_real);
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> u(0.f, 1.f);
    for (auto& v : h_in) v = u(rng);
    const size_t bytes = plan.get_alloc_bytes();
    float* a = plan.mem_alloc<float>(bytes);
    auto* b = plan.mem_alloc<std::complex<float>>(bytes);
    float* c = plan.mem_alloc<float>(bytes);
    cudaStream_t stream = static_cast<cudaStream_t>(plan.get_stream());
    CUDA_OK(cudaMemcpy(a, h_in.data(), n_real * sizeof(float), cudaMemcpyHostToDevice));
    EXPECT(plan.execute(a, b, Execute::FORWARD) == Error::SUCCESS);
    CUDA_OK(cudaStreamSynchronize(stream));
    // DC term = sum of the input (Y pencil element (ky = 0, kx = 0) is the first one)
    std::complex<float> dc;
    CUDA_OK(cudaMemcpy(&dc, b, sizeof(dc), cudaMemcpyDeviceToHost));
    double sum = 0;
    for (float v : h_in) sum += v;
    EXPECT(std::abs(dc.real() - sum) <= 1e-5 * sum && std::abs(dc.imag()) <= 1e-5 * sum);
    EXPECT(plan.execute(b, c, Execute::BACKWARD) == Error::SUCCESS);
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaMemcpy(h_back.data(), c, n_real * sizeof(float), cudaMemcpyDeviceToHost));
    const double scale = 1.0 / (double)n_real;
    const double bound = 5.0 * std::log2((double)n_real) * 2.0 * 1.1920929e-07;
    double worst = 0;
    for (size_t i = 0; i < n_real; ++i) worst = std::max(worst, std::abs(h_back[i] * scale - h_in[i]));
    EXPECT(worst <= bound);
    // third call with the same buffers replays the captured CUDA graph: same bits
    std::vector<std::complex<float>> s1(alloc), s2(alloc);
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_OK(cudaMemcpy(a, h_in.data(), n_real * sizeof(float), cudaMemcpyHostToDevice));
        EXPECT(plan.execute(a, b, Execute::FORWARD) == Error::SUCCESS);
        CUDA_OK(cudaStreamSynchronize(stream));
        CUDA_OK(cudaMemcpy((rep == 0 ? s1 : s2).data(), b, (size_t)40 * 34 * sizeof(std::complex<float>), cudaMemcpyDeviceToHost));
    }
    EXPECT(std::memcmp(s1.data(), s2.data(), (size_t)40 * 34 * sizeof(std::complex<float>)) == 0);
    EXPECT(plan.mem_free(a) == Error::SUCCESS && plan.mem_free(b) == Error::SUCCESS && plan.mem_free(c) == Error::SUCCESS);
}

static void user_pencil_and_r2r() {
    // one rank describing the whole box as a "pencil" + an R2R transpose-only plan in single precision
    EXPECT(set_config(Config()) == Error::SUCCESS);
    const std::vector<int32_t> starts = {0, 0, 0}, counts = {18, 33, 155};
    Pencil box(starts, counts);
    EXPECT(box.get_size() == (size_t)18 * 33 * 155 && box.get_ndims() == 3);
    PlanR2R plan(box, Precision::SINGLE);
    EXPECT(plan.get_dims() == counts && plan.get_element_size() == 4);
    EXPECT(plan.get_z_slab_enabled());  // one rank: Nz / 1 >= 32
    const size_t n = box.get_size();
    std::vector<float> h(n), back(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)i;  // in(i) = i like src/tests/test_host_kernels.F90:35-37
    float* a = plan.mem_alloc<float>(plan.get_alloc_bytes());
    float* b = plan.mem_alloc<float>(plan.get_alloc_bytes());
    CUDA_OK(cudaMemcpy(a, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    EXPECT(plan.transpose(a, b, Transpose::X_TO_Z) == Error::SUCCESS);
    EXPECT(plan.transpose(b, a, Transpose::Z_TO_X) == Error::SUCCESS);
    CUDA_OK(cudaStreamSynchronize(static_cast<cudaStream_t>(plan.get_stream())));
    CUDA_OK(cudaMemcpy(back.data(), a, n * sizeof(float), cudaMemcpyDeviceToHost));
    EXPECT(std::memcmp(back.data(), h.data(), n * sizeof(float)) == 0);
    bool thrown = false;
    try {
        Pencil nothing;
        (void)nothing.get_size();
    } catch (const Exception& e) {
        thrown = e.get_error_code() == Error::PENCIL_NOT_INITIALIZED;
    }
    EXPECT(thrown);
    EXPECT(plan.mem_free(a) == Error::SUCCESS && plan.mem_free(b) == Error::SUCCESS);
}

int main() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        std::fprintf(stderr, "api_gpu needs a CUDA device\n");
        return 2;
    }
    EXPECT(Version::get() == Version::CODE && Version::get(3, 2, 0) == Version::CODE);
    EXPECT(!get_error_string(Error::INVALID_AUX).empty() && get_backend_pipelined(Backend::NCCL_PIPELINED));
    try {
        transpose_only_3d();
        r2c_2d_float_cufft();
        user_pencil_and_r2r();
    } catch (const Exception& e) {
        std::fprintf(stderr, "unexpected %s\n", e.what());
        ++failures;
    }
    if (failures) {
        std::fprintf(stderr, "api_gpu: %d check(s) failed\n", failures);
        return 1;
    }
    std::printf("api_gpu OK\n");
    return 0;
}
