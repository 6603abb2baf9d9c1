// nvlink_probe: what does one B200 push through NVLink from a KERNEL, and with which access form?
// Single process, two devices (0 = sender, 1 = receiver), peer access enabled; both directions can
// run at once (--bidir).  Every variant moves `bytes` from a local source to a peer destination laid
// out as rows of `run` contiguous bytes at a destination pitch of `pitch` bytes -- the shape of the
// fused Y<->Z transposition of the 512^3 cycle at 8 GPUs (1 KB runs, 8 KB pitch).
//
//   st128      plain 16-byte stores to the peer (what transpose_tiles_kernel does today)
//   bulk       rows staged in shared memory, written with cp.async.bulk.global.shared::cta (TMA 1-D)
//   tma2tma    cp.async.bulk global->shared (mbarrier) + cp.async.bulk shared->peer: no SM data path
//   pull128    receiver side: 16-byte loads from the peer, local stores
//   memcpy     cudaMemcpyPeerAsync (copy engine), contiguous
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o nvlink_probe nvlink_probe.cu
// Prints one JSON line per (variant, run, direction).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- st128: thread t moves 16 B; a row of `run` bytes goes to dst + row * pitch ----------------
template <int UNROLL>
__global__ void __launch_bounds__(256) k_st128(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n16,
                                               int run16, long long pitch16) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = src[i + u * stride];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long j = i + u * stride;
            const long long row = j / run16;
            dst[row * pitch16 + (j - row * run16)] = v[u];
        }
    }
    for (; i < n16; i += stride) {
        const long long row = i / run16;
        dst[row * pitch16 + (i - row * run16)] = src[i];
    }
}

// ---- pull128: same index map, but the SOURCE is the peer and the destination is local ----------
template <int UNROLL>
__global__ void __launch_bounds__(256) k_pull128(const uint4* __restrict__ peer_src, uint4* __restrict__ dst, long long n16,
                                                 int run16, long long pitch16) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = peer_src[i + u * stride];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long j = i + u * stride;
            const long long row = j / run16;
            dst[row * pitch16 + (j - row * run16)] = v[u];
        }
    }
    for (; i < n16; i += stride) {
        const long long row = i / run16;
        dst[row * pitch16 + (i - row * run16)] = peer_src[i];
    }
}

// ---- bulk: 256 threads fill a stage of ROWS rows with 16-byte loads, one thread per row issues the TMA store
template <int STAGES>
__global__ void __launch_bounds__(256) k_bulk(const uint4* __restrict__ src, unsigned char* __restrict__ dst, long long nrows,
                                              int run, long long pitch, int rows_per_stage) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int stage_bytes = rows_per_stage * run;
    const int run16 = run / 16;
    const int stage16 = stage_bytes / 16;
    int s = 0;
    for (long long r0 = (long long)blockIdx.x * rows_per_stage; r0 < nrows; r0 += (long long)gridDim.x * rows_per_stage) {
        unsigned char* buf = smem + (size_t)s * stage_bytes;
        // the bulk store that last read this stage must have finished reading shared memory
        if (threadIdx.x < rows_per_stage) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory");
        __syncthreads();
        const long long rows = nrows - r0 < rows_per_stage ? nrows - r0 : rows_per_stage;
        const uint4* g = src + r0 * run16;
        for (int i = threadIdx.x; i < stage16; i += blockDim.x)
            if (i < rows * run16) reinterpret_cast<uint4*>(buf)[i] = g[i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < rows_per_stage) {
            if (threadIdx.x < rows) {
                unsigned char* d = dst + (r0 + threadIdx.x) * pitch;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d),
                             "r"(smem_u32(buf + (size_t)threadIdx.x * run)), "r"(run)
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        s = (s + 1) % STAGES;
    }
    if (threadIdx.x < rows_per_stage) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- tma2tma: one thread per CTA drives a STAGES-deep ring: bulk load of a stage (contiguous in the
// source) signalled on an mbarrier, then one bulk store per row to the peer --------------------------------
template <int STAGES>
__global__ void __launch_bounds__(32) k_tma2tma(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst,
                                                long long nrows, int run, long long pitch, int rows_per_stage) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar[STAGES];
    if (threadIdx.x != 0) return;
    const int stage_bytes = rows_per_stage * run;
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long step = (long long)gridDim.x * rows_per_stage;
    long long r_load = (long long)blockIdx.x * rows_per_stage, r_store = r_load;
    int s_load = 0, s_store = 0;
    unsigned phase_bits = 0;
    int inflight = 0;
    auto issue_load = [&]() {
        const long long rows = nrows - r_load < rows_per_stage ? nrows - r_load : rows_per_stage;
        const unsigned bytes = (unsigned)(rows * run);
        // the store that last read this stage must be done with shared memory
        // (loads run STAGES - 2 ahead of the stores, so all but the newest store group must have been read)
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s_load])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + (size_t)s_load * stage_bytes)),
                     "l"(src + r_load * run), "r"(bytes), "r"(smem_u32(&bar[s_load]))
                     : "memory");
        r_load += step;
        s_load = (s_load + 1) % STAGES;
        ++inflight;
    };
    while (inflight < STAGES - 2 && r_load < nrows) issue_load();
    while (r_store < nrows) {
        if (r_load < nrows) issue_load();
        // wait for the stage to land
        const unsigned parity = (phase_bits >> s_store) & 1u;
        unsigned done = 0;
        while (!done)
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(smem_u32(&bar[s_store])), "r"(parity)
                : "memory");
        phase_bits ^= 1u << s_store;
        const long long rows = nrows - r_store < rows_per_stage ? nrows - r_store : rows_per_stage;
        for (int r = 0; r < rows; ++r)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (r_store + r) * pitch),
                         "r"(smem_u32(smem + (size_t)s_store * stage_bytes + (size_t)r * run)), "r"(run)
                         : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        r_store += step;
        s_store = (s_store + 1) % STAGES;
        --inflight;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- local HBM load to run beside a link variant: plain 16-byte copy inside one device -----------------
__global__ void __launch_bounds__(256) k_local_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n16) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
}

struct Side {
    int dev;
    unsigned char *src, *dst;  // src is local to dev, dst lives on the OTHER device
    cudaStream_t stream, bg_stream;
    cudaEvent_t e0, e1, b0, b1;
    unsigned char *bg_src, *bg_dst;  // local buffers of the background HBM load
};

int main(int argc, char** argv) {
    size_t bytes = 256ull << 20;
    int iters = 10;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--mb") && i + 1 < argc) bytes = (size_t)atoll(argv[++i]) << 20;
        if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atoi(argv[++i]);
    }
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) {
        printf("{\"error\": \"needs 2 devices, found %d\"}\n", ndev);
        return 0;
    }
    const size_t max_pitch_factor = 8;
    const size_t bg_bytes = 1ull << 30;
    Side side[2];
    unsigned char* bufs[2][2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&bufs[d][0], bytes));
        CK(cudaMalloc(&bufs[d][1], bytes * max_pitch_factor));
        CK(cudaMemset(bufs[d][0], d + 1, bytes));
    }
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        side[d].dev = d;
        side[d].src = bufs[d][0];
        side[d].dst = bufs[1 - d][1];
        CK(cudaStreamCreateWithFlags(&side[d].stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&side[d].bg_stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&side[d].e0));
        CK(cudaEventCreate(&side[d].e1));
        CK(cudaEventCreate(&side[d].b0));
        CK(cudaEventCreate(&side[d].b1));
        CK(cudaMalloc(&side[d].bg_src, bg_bytes));
        CK(cudaMalloc(&side[d].bg_dst, bg_bytes));
        CK(cudaMemset(side[d].bg_src, 7, bg_bytes));
    }
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

    struct Var {
        std::string name;
        int run;        // contiguous destination bytes
        int pitch_mul;  // destination pitch = run * pitch_mul
        int param;      // rows per stage / CTAs per SM
    };
    std::vector<Var> vars;
    for (int run : {512, 1024, 2048, 4096}) {
        for (int pm : {1, 8}) {
            vars.push_back({"st128", run, pm, 8});
            vars.push_back({"pull128", run, pm, 8});
            vars.push_back({"bulk", run, pm, 16384 / run});
            vars.push_back({"bulk", run, pm, 32768 / run});
            vars.push_back({"tma2tma", run, pm, 16384 / run});
            vars.push_back({"tma2tma", run, pm, 32768 / run});
        }
    }
    vars.push_back({"memcpy", 0, 1, 0});
    for (int run : {512, 1024, 2048, 4096}) {
        vars.push_back({"memcpy2d", run, 8, 0});
        vars.push_back({"memcpy3d", run, 8, 512});  // planes of 512 rows, like one y-plane of the 512^3 Z pencil
    }

    auto launch = [&](const Var& v, Side& s) {
        CK(cudaSetDevice(s.dev));
        const long long n16 = (long long)(bytes / 16);
        if (v.name == "memcpy") {
            CK(cudaMemcpyPeerAsync(s.dst, 1 - s.dev, s.src, s.dev, bytes, s.stream));
            return;
        }
        const long long nrows = (long long)(bytes / v.run);
        const long long pitch = (long long)v.run * v.pitch_mul;
        if (v.name == "memcpy2d") {  // copy engine, rows of `run` bytes at the destination pitch
            CK(cudaMemcpy2DAsync(s.dst, (size_t)pitch, s.src, (size_t)v.run, (size_t)v.run, (size_t)nrows,
                                 cudaMemcpyDeviceToDevice, s.stream));
            return;
        }
        if (v.name == "memcpy3d") {
            cudaMemcpy3DPeerParms p3{};
            const size_t rows = (size_t)v.param, planes = (size_t)nrows / rows;
            p3.srcDevice = s.dev, p3.dstDevice = 1 - s.dev;
            p3.srcPtr = make_cudaPitchedPtr(s.src, (size_t)v.run, (size_t)v.run, rows);
            p3.dstPtr = make_cudaPitchedPtr(s.dst, (size_t)pitch, (size_t)v.run, rows);
            p3.extent = make_cudaExtent((size_t)v.run, rows, planes);
            CK(cudaMemcpy3DPeerAsync(&p3, s.stream));
            return;
        }
        if (v.name == "st128") {
            k_st128<8><<<sms * v.param, 256, 0, s.stream>>>((const uint4*)s.src, (uint4*)s.dst, n16, v.run / 16, pitch / 16);
        } else if (v.name == "pull128") {
            // receiver-side kernel: runs on s.dev, reads the OTHER device's source, writes its own buffer
            k_pull128<8><<<sms * v.param, 256, 0, s.stream>>>((const uint4*)bufs[1 - s.dev][0], (uint4*)bufs[s.dev][1], n16,
                                                              v.run / 16, pitch / 16);
        } else if (v.name == "bulk") {
            constexpr int ST = 3;
            const size_t smem = (size_t)ST * v.param * v.run;
            static bool set = false;
            if (!set) {
                for (int d = 0; d < 2; ++d) {
                    CK(cudaSetDevice(d));
                    CK(cudaFuncSetAttribute(k_bulk<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                }
                CK(cudaSetDevice(s.dev));
                set = true;
            }
            k_bulk<ST><<<sms * 2, 256, smem, s.stream>>>((const uint4*)s.src, s.dst, nrows, v.run, pitch, v.param);
        } else if (v.name == "tma2tma") {
            constexpr int ST = 4;
            const size_t smem = (size_t)ST * v.param * v.run;
            static bool set = false;
            if (!set) {
                for (int d = 0; d < 2; ++d) {
                    CK(cudaSetDevice(d));
                    CK(cudaFuncSetAttribute(k_tma2tma<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                }
                CK(cudaSetDevice(s.dev));
                set = true;
            }
            k_tma2tma<ST><<<sms, 32, smem, s.stream>>>(s.src, s.dst, nrows, v.run, pitch, v.param);
        }
        CK(cudaGetLastError());
    };

    for (int pass = 0; pass < 3; ++pass) {
        const int bidir = pass >= 1 ? 1 : 0;
        const bool with_bg = pass == 2;  // both directions busy AND a local HBM copy running on every device
        for (const Var& v : vars) {
            if (with_bg && !(v.name == "memcpy" || ((v.name == "st128" || v.name == "memcpy2d" || v.name == "memcpy3d") && v.run == 1024 && v.pitch_mul == 8)))
                continue;
            const int nsides = bidir ? 2 : 1;
            for (int w = 0; w < 2; ++w)
                for (int d = 0; d < nsides; ++d) launch(v, side[d]);
            for (int d = 0; d < 2; ++d) {
                CK(cudaSetDevice(d));
                CK(cudaDeviceSynchronize());
            }
            const int bg_iters = 3 * iters;
            if (with_bg)
                for (int d = 0; d < 2; ++d) {
                    CK(cudaSetDevice(d));
                    CK(cudaEventRecord(side[d].b0, side[d].bg_stream));
                    for (int it = 0; it < bg_iters; ++it)
                        k_local_copy<<<sms * 8, 256, 0, side[d].bg_stream>>>((const uint4*)side[d].bg_src, (uint4*)side[d].bg_dst,
                                                                              (long long)(bg_bytes / 16));
                    CK(cudaEventRecord(side[d].b1, side[d].bg_stream));
                }
            for (int d = 0; d < nsides; ++d) {
                CK(cudaSetDevice(d));
                CK(cudaEventRecord(side[d].e0, side[d].stream));
            }
            for (int it = 0; it < iters; ++it)
                for (int d = 0; d < nsides; ++d) launch(v, side[d]);
            float worst = 0.f;
            for (int d = 0; d < nsides; ++d) {
                CK(cudaSetDevice(d));
                CK(cudaEventRecord(side[d].e1, side[d].stream));
            }
            for (int d = 0; d < nsides; ++d) {
                CK(cudaSetDevice(d));
                CK(cudaEventSynchronize(side[d].e1));
                float ms = 0.f;
                CK(cudaEventElapsedTime(&ms, side[d].e0, side[d].e1));
                if (ms > worst) worst = ms;
            }
            const double gbs = (double)bytes * iters / (worst * 1e-3) / 1e9;
            double bg_gbs = 0.0;
            if (with_bg) {
                float bg_worst = 0.f;
                for (int d = 0; d < 2; ++d) {
                    CK(cudaSetDevice(d));
                    CK(cudaEventSynchronize(side[d].b1));
                    float ms = 0.f;
                    CK(cudaEventElapsedTime(&ms, side[d].b0, side[d].b1));
                    if (ms > bg_worst) bg_worst = ms;
                }
                // read + write bytes of the local copy over ITS whole run (part of it alone, after the link variant ended)
                bg_gbs = 2.0 * (double)bg_bytes * bg_iters / (bg_worst * 1e-3) / 1e9;
            }
            printf("{\"variant\": \"%s\", \"run_bytes\": %d, \"pitch_mul\": %d, \"param\": %d, \"bidir\": %d, \"mb\": %zu, "
                   "\"ms\": %.4f, \"GBps_per_direction\": %.1f, \"with_local_hbm_load\": %d, \"local_copy_GBps_over_its_run\": %.1f}\n",
                   v.name.c_str(), v.run, v.pitch_mul, v.param, bidir, bytes >> 20, worst / iters, gbs, with_bg ? 1 : 0, bg_gbs);
            fflush(stdout);
        }
    }
    // spot check: the last variant copied the pattern d+1 of the sender
    for (int d = 0; d < 2; ++d) {
        unsigned char h[16];
        CK(cudaSetDevice(d));
        CK(cudaMemcpy(h, bufs[d][1], 16, cudaMemcpyDeviceToHost));
        if (h[0] != (unsigned char)(2 - d)) fprintf(stderr, "warning: device %d destination holds %d\n", d, (int)h[0]);
    }
    return 0;
}
