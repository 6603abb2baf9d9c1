// Peer-memory registry + device barrier (see peer.h).
#include "peer.h"

#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include "errors.h"

namespace dtfftb {

namespace {

struct RankId {
    char host[64];
    int device;
    int pid;
};

struct IpcMsg {
    cudaIpcMemHandle_t handle;
    unsigned long long offset;  // of the registered pointer inside the exported allocation
    unsigned long long bytes;
    int ok;
};

// Driver entry point fetched at run time: the library must load on boxes without libcuda.
typedef int (*cuMemGetAddressRange_t)(unsigned long long* pbase, size_t* psize, unsigned long long dptr);

void* allocation_base(void* ptr) {
    static cuMemGetAddressRange_t fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<cuMemGetAddressRange_t>(f);
        cudaGetLastError();
    }
    if (!fn) return ptr;
    unsigned long long base = 0;
    size_t size = 0;
    if (fn(&base, &size, (unsigned long long)(uintptr_t)ptr) != 0) return ptr;
    return reinterpret_cast<void*>((uintptr_t)base);
}

typedef int (*cuPointerGetAttribute_t)(void* data, int attribute, unsigned long long ptr);

}  // namespace

// CU_POINTER_ATTRIBUTE_BUFFER_ID: unique per allocation for the life of the process, so a buffer that was
// freed and re-allocated at the same address is told apart from the one whose peer mappings are cached.
unsigned long long buffer_id(const void* ptr) {
    static cuPointerGetAttribute_t fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<cuPointerGetAttribute_t>(f);
        cudaGetLastError();
    }
    if (!fn || !ptr) return 0;
    unsigned long long id = 0;
    if (fn(&id, 7 /* CU_POINTER_ATTRIBUTE_BUFFER_ID */, (unsigned long long)(uintptr_t)ptr) != 0) return 0;
    return id;
}

namespace {

// An allocation can be opened only once per process: cache by handle bytes.
std::map<std::string, std::pair<void*, int>>& ipc_cache() {
    static std::map<std::string, std::pair<void*, int>> c;
    return c;
}

void* ipc_open(const cudaIpcMemHandle_t& h) {
    std::string key(reinterpret_cast<const char*>(&h), sizeof(h));
    auto& c = ipc_cache();
    auto it = c.find(key);
    if (it != c.end()) {
        it->second.second++;
        return it->second.first;
    }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    c[key] = std::make_pair(p, 1);
    return p;
}

void ipc_close(void* p) {
    if (!p) return;
    auto& c = ipc_cache();
    for (auto it = c.begin(); it != c.end(); ++it)
        if (it->second.first == p) {
            if (--it->second.second == 0) {
                cudaIpcCloseMemHandle(p);
                cudaGetLastError();
                c.erase(it);
            }
            return;
        }
}

// One block; thread t handles member t.  The epoch lives in device memory and is advanced by the
// kernel itself, so the launch carries no per-call argument and can be replayed from a CUDA graph.
__global__ void peer_barrier_kernel(uint64_t* const* __restrict__ peer_flags, uint64_t* __restrict__ my_flags,
                                    const int* __restrict__ members, int n, int me, size_t row,
                                    uint64_t* __restrict__ epoch_counter, long long timeout_cycles,
                                    uint64_t* __restrict__ err, uint64_t* __restrict__ err_host) {
    __shared__ uint64_t s_epoch;
    const int t = threadIdx.x;
    if (t == 0) {
        s_epoch = *epoch_counter + 1;
        *epoch_counter = s_epoch;
    }
    __syncthreads();
    if (t >= n) return;
    const uint64_t epoch = s_epoch;
    const int peer = members[t];
    uint64_t* remote = peer_flags[t] + row + me;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(epoch) : "memory");
    const uint64_t* mine = my_flags + row + peer;
    const long long t0 = clock64();
    uint64_t v;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= epoch) break;
        if (clock64() - t0 > timeout_cycles) {  // fatal: later kernels store nothing, the next API call fails
            *reinterpret_cast<volatile uint64_t*>(err) = 1;
            *reinterpret_cast<volatile uint64_t*>(err_host) = 1;
            __threadfence_system();
            break;
        }
        __nanosleep(64);
    }
}

__global__ void peer_epoch_kernel(uint64_t* __restrict__ epoch_counter) { *epoch_counter += 1; }

// Enqueued behind the copy that carries the data: the copy has completed when this runs.
__global__ void peer_signal_kernel(uint64_t* __restrict__ remote_flag, const uint64_t* __restrict__ epoch_counter,
                                   uint64_t slices, uint64_t slice) {
    const uint64_t epoch = *reinterpret_cast<const volatile uint64_t*>(epoch_counter) * slices + slice + 1;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote_flag), "l"(epoch) : "memory");
}

__global__ void peer_wait_kernel(const uint64_t* __restrict__ my_flag, const uint64_t* __restrict__ epoch_counter,
                                 uint64_t slices, uint64_t slice, long long timeout_cycles, uint64_t* __restrict__ err,
                                 uint64_t* __restrict__ err_host) {
    const uint64_t epoch = *reinterpret_cast<const volatile uint64_t*>(epoch_counter) * slices + slice + 1;
    const long long t0 = clock64();
    uint64_t v;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(my_flag) : "memory");
        if (v >= epoch) break;
        if (clock64() - t0 > timeout_cycles) {  // fatal, see peer_barrier_kernel
            *reinterpret_cast<volatile uint64_t*>(err) = 1;
            *reinterpret_cast<volatile uint64_t*>(err_host) = 1;
            __threadfence_system();
            break;
        }
        __nanosleep(64);
    }
}

}  // namespace

int PeerRegistry::init(const Comm& world) {
    destroy();
    world_ = world;
    inited_ = true;
    available_ = false;
    shared_device_ = false;
    const int P = world_.size();
    RankId me{};
    gethostname(me.host, sizeof(me.host) - 1);
    cudaError_t ce = cudaGetDevice(&me.device);
    if (ce != cudaSuccess) return cuda_error(ce);
    me.pid = (int)getpid();
    std::vector<RankId> all;
    if (world_.allgather_v(me, all)) return DTFFTB_ERROR_INTERNAL;
    int ok = 1;
    why_ = "";
    for (int r = 0; r < P && ok; ++r) {
        if (std::strncmp(all[r].host, me.host, sizeof(me.host)) != 0) ok = 0, why_ = "ranks span several hosts";
        if (r != world_.rank() && all[r].pid == me.pid) ok = 0, why_ = "several ranks in one process";
        if (r != world_.rank() && ok) {
            int can = 0;
            if (all[r].device == me.device) {
                // DTFFTB_ALLOW_SHARED_DEVICE=1 (tests on a 1-GPU box): ranks time-slice one device; cudaIpc and the
                // spin barriers work (a spinning kernel is preempted at the end of its time slice), only slowly
                const char* e = getenv("DTFFTB_ALLOW_SHARED_DEVICE");
                if (e && atoi(e))
                    shared_device_ = true;
                else
                    ok = 0, why_ = "two ranks share a device";
            } else if (cudaDeviceCanAccessPeer(&can, me.device, all[r].device) != cudaSuccess || !can)
                ok = 0, why_ = "no peer access between devices";
        }
    }
    cudaGetLastError();
    if (const char* e = getenv("DTFFTB_DISABLE_P2P"))
        if (atoi(e)) ok = 0, why_ = "disabled by DTFFTB_DISABLE_P2P";
    shared_device_ = world_.sum(shared_device_ ? 1 : 0) > 0;  // every rank must know (Plan::create skips NCCL then)
    if (world_.sum(ok) != P) {
        if (!*why_) why_ = "a peer reported no access";
        return DTFFT_SUCCESS;
    }
    if (P == 1) {
        available_ = true;
        return DTFFT_SUCCESS;
    }
    {  // time-out of the device barriers: DTFFTB_PEER_TIMEOUT_MS milliseconds at the SM clock
        long long ms = 20000;
        if (const char* e = getenv("DTFFTB_PEER_TIMEOUT_MS")) ms = std::max(1ll, atoll(e));
        int khz = 2000000;
        if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, me.device) != cudaSuccess || khz <= 0) khz = 2000000;
        cudaGetLastError();
        timeout_cycles_ = ms * (long long)khz;
    }
    ce = cudaHostAlloc(reinterpret_cast<void**>(&h_err_), sizeof(uint64_t), cudaHostAllocMapped);
    if (ce != cudaSuccess) return cuda_error(ce);
    *h_err_ = 0;
    ce = cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_err_host_), h_err_, 0);
    if (ce != cudaSuccess) return cuda_error(ce);
    const size_t n = (size_t)kChannels * P + 8;
    ce = cudaMalloc(&flags_, n * sizeof(uint64_t));
    if (ce != cudaSuccess) return cuda_error(ce);
    ce = cudaMemset(flags_, 0, n * sizeof(uint64_t));
    if (ce != cudaSuccess) return cuda_error(ce);
    available_ = true;
    int rc = register_buffer(flags_, n * sizeof(uint64_t), &flags_slot_);
    if (rc != DTFFT_SUCCESS || flags_slot_ < 0) {
        available_ = false;
        why_ = "cudaIpc exchange of the flags buffer failed";
        return rc;
    }
    return DTFFT_SUCCESS;
}

int PeerRegistry::register_buffer(void* ptr, size_t bytes, int* slot_out) {
    if (slot_out) *slot_out = -1;
    if (!inited_ || !available_) return DTFFT_SUCCESS;
    const int P = world_.size();
    Slot s;
    s.local = ptr;
    s.bytes = bytes;
    s.mapped.assign((size_t)P, nullptr);
    s.opened.assign((size_t)P, nullptr);
    s.mapped[(size_t)world_.rank()] = ptr;
    if (P > 1) {
        IpcMsg mine{};
        void* base = allocation_base(ptr);
        mine.offset = (unsigned long long)((char*)ptr - (char*)base);
        mine.bytes = bytes;
        mine.ok = cudaIpcGetMemHandle(&mine.handle, base) == cudaSuccess ? 1 : 0;
        cudaGetLastError();
        std::vector<IpcMsg> all;
        if (world_.allgather_v(mine, all)) return DTFFTB_ERROR_INTERNAL;
        int ok = 1;
        for (int r = 0; r < P; ++r) ok &= all[r].ok;
        for (int r = 0; r < P && ok; ++r) {
            if (r == world_.rank()) continue;
            void* p = ipc_open(all[r].handle);
            if (!p) {
                ok = 0;
                break;
            }
            s.opened[(size_t)r] = p;
            s.mapped[(size_t)r] = (char*)p + all[r].offset;
        }
        // every rank must agree, otherwise nobody uses the slot
        if (world_.sum(ok) != P) {
            for (void* p : s.opened)
                if (p) ipc_close(p);
            cudaGetLastError();
            return DTFFT_SUCCESS;  // not registered; callers see resolve() == false
        }
    }
    s.live = true;
    int slot = -1;
    for (size_t i = 0; i < slots_.size(); ++i)
        if (!slots_[i].live) {
            slot = (int)i;
            break;
        }
    if (slot < 0) {
        slots_.push_back(Slot{});
        slot = (int)slots_.size() - 1;
    }
    slots_[(size_t)slot] = s;
    if (slot_out) *slot_out = slot;
    return DTFFT_SUCCESS;
}

int PeerRegistry::publish(void* ptr, size_t bytes, std::vector<void*>* mapped, std::vector<void*>* opened, bool* ok) {
    *ok = false;
    mapped->clear();
    opened->clear();
    if (!inited_ || !available_) return DTFFT_SUCCESS;
    const int P = world_.size();
    mapped->assign((size_t)P, nullptr);
    (*mapped)[(size_t)world_.rank()] = ptr;
    if (P == 1) {
        *ok = true;
        return DTFFT_SUCCESS;
    }
    IpcMsg mine{};
    void* base = allocation_base(ptr);
    mine.offset = (unsigned long long)((char*)ptr - (char*)base);
    mine.bytes = bytes;
    mine.ok = cudaIpcGetMemHandle(&mine.handle, base) == cudaSuccess ? 1 : 0;
    cudaGetLastError();
    std::vector<IpcMsg> all;
    if (world_.allgather_v(mine, all)) return DTFFTB_ERROR_COMM;
    int good = 1;
    for (int r = 0; r < P; ++r) good &= all[r].ok;
    opened->assign((size_t)P, nullptr);
    for (int r = 0; r < P && good; ++r) {
        if (r == world_.rank()) continue;
        void* p = ipc_open(all[r].handle);
        if (!p) {
            good = 0;
            break;
        }
        (*opened)[(size_t)r] = p;
        (*mapped)[(size_t)r] = (char*)p + all[r].offset;
    }
    if (world_.sum(good) != P) {  // every rank must agree, otherwise nobody uses the mapping
        release(opened);
        mapped->clear();
        return DTFFT_SUCCESS;
    }
    *ok = true;
    return DTFFT_SUCCESS;
}

void PeerRegistry::release(std::vector<void*>* opened) {
    for (void* p : *opened)
        if (p) ipc_close(p);
    cudaGetLastError();
    opened->clear();
}

int PeerRegistry::unregister_buffer(void* ptr) {
    if (!inited_) return DTFFT_SUCCESS;
    for (auto& s : slots_) {
        if (!s.live || s.local != ptr) continue;
        // peers may still be storing into this buffer: wait for everybody
        cudaDeviceSynchronize();
        world_.barrier();
        for (void* p : s.opened)
            if (p) ipc_close(p);
        cudaGetLastError();
        world_.barrier();
        s = Slot{};
        return DTFFT_SUCCESS;
    }
    return DTFFT_SUCCESS;
}

bool PeerRegistry::resolve(const void* ptr, size_t bytes, int* slot, size_t* offset) const {
    const char* p = (const char*)ptr;
    for (size_t i = 0; i < slots_.size(); ++i) {
        const Slot& s = slots_[i];
        if (!s.live) continue;
        const char* b = (const char*)s.local;
        if (p >= b && p + bytes <= b + s.bytes) {
            *slot = (int)i;
            *offset = (size_t)(p - b);
            return true;
        }
    }
    return false;
}

void* PeerRegistry::peer_ptr(int r, int slot, size_t offset) const {
    return (char*)slots_[(size_t)slot].mapped[(size_t)r] + offset;
}

PeerRegistry::Group* PeerRegistry::group_for(const std::vector<int>& members, int channel, int* rc) {
    *rc = DTFFT_SUCCESS;
    const int n = (int)members.size();
    auto key = std::make_pair(channel, members);
    auto it = groups_.find(key);
    if (it == groups_.end()) {
        Group g;
        g.n = n;
        std::vector<uint64_t*> bases((size_t)n);
        for (int t = 0; t < n; ++t) bases[(size_t)t] = (uint64_t*)peer_ptr(members[(size_t)t], flags_slot_, 0);
        cudaError_t ce = cudaMalloc(&g.d_peer_flags, n * sizeof(uint64_t*));
        if (ce == cudaSuccess) ce = cudaMalloc(&g.d_members, n * sizeof(int));
        if (ce == cudaSuccess) ce = cudaMalloc(&g.d_epoch, sizeof(uint64_t));
        if (ce != cudaSuccess) {
            *rc = cuda_error(ce);
            return nullptr;
        }
        cudaMemset(g.d_epoch, 0, sizeof(uint64_t));
        cudaMemcpy(g.d_peer_flags, bases.data(), n * sizeof(uint64_t*), cudaMemcpyHostToDevice);
        cudaMemcpy(g.d_members, members.data(), n * sizeof(int), cudaMemcpyHostToDevice);
        it = groups_.emplace(key, g).first;
    }
    return &it->second;
}

int PeerRegistry::barrier(const std::vector<int>& members, int channel, cudaStream_t stream) {
    if (!available_) return DTFFTB_ERROR_INTERNAL;
    const int n = (int)members.size();
    if (n <= 1) return DTFFT_SUCCESS;
    if (channel < 0 || channel >= kChannels || n > 1024) return DTFFTB_ERROR_INTERNAL;
    int grc = 0;
    Group* gp = group_for(members, channel, &grc);
    if (!gp) return grc;
    Group& g = *gp;
    const int P = world_.size();
    uint64_t* err = flags_ + (size_t)kChannels * P;
    // a missing or late peer turns into a sticky, fatal error instead of a hung GPU
    const int threads = ((n + 31) / 32) * 32;
    peer_barrier_kernel<<<1, threads, 0, stream>>>(g.d_peer_flags, flags_, g.d_members, n, world_.rank(),
                                                   (size_t)channel * P, g.d_epoch, timeout_cycles_, err, d_err_host_);
    cudaError_t ce = cudaGetLastError();
    return ce == cudaSuccess ? DTFFT_SUCCESS : cuda_error(ce);
}

int PeerRegistry::advance_epoch(const std::vector<int>& members, int channel, cudaStream_t stream) {
    if (!available_) return DTFFTB_ERROR_INTERNAL;
    if (members.size() <= 1) return DTFFT_SUCCESS;
    if (channel < 0 || channel >= kChannels) return DTFFTB_ERROR_INTERNAL;
    int grc = 0;
    Group* g = group_for(members, channel, &grc);
    if (!g) return grc;
    peer_epoch_kernel<<<1, 1, 0, stream>>>(g->d_epoch);
    cudaError_t ce = cudaGetLastError();
    return ce == cudaSuccess ? DTFFT_SUCCESS : cuda_error(ce);
}

int PeerRegistry::signal(const std::vector<int>& members, int channel, int member_index, int slice, cudaStream_t stream) {
    if (!available_ || channel < 0 || channel >= kChannels || slice < 0 || slice >= kMaxSlices) return DTFFTB_ERROR_INTERNAL;
    if (member_index < 0 || member_index >= (int)members.size()) return DTFFTB_ERROR_INTERNAL;
    int grc = 0;
    Group* g = group_for(members, channel, &grc);
    if (!g) return grc;
    const int P = world_.size();
    uint64_t* remote = (uint64_t*)peer_ptr(members[(size_t)member_index], flags_slot_, 0) + (size_t)channel * P + world_.rank();
    peer_signal_kernel<<<1, 1, 0, stream>>>(remote, g->d_epoch, (uint64_t)kMaxSlices, (uint64_t)slice);
    cudaError_t ce = cudaGetLastError();
    return ce == cudaSuccess ? DTFFT_SUCCESS : cuda_error(ce);
}

int PeerRegistry::wait(const std::vector<int>& members, int channel, int member_index, int slice, cudaStream_t stream) {
    if (!available_ || channel < 0 || channel >= kChannels || slice < 0 || slice >= kMaxSlices) return DTFFTB_ERROR_INTERNAL;
    if (member_index < 0 || member_index >= (int)members.size()) return DTFFTB_ERROR_INTERNAL;
    int grc = 0;
    Group* g = group_for(members, channel, &grc);
    if (!g) return grc;
    const int P = world_.size();
    const uint64_t* mine = flags_ + (size_t)channel * P + members[(size_t)member_index];
    peer_wait_kernel<<<1, 1, 0, stream>>>(mine, g->d_epoch, (uint64_t)kMaxSlices, (uint64_t)slice, timeout_cycles_,
                                          flags_ + (size_t)kChannels * P, d_err_host_);
    cudaError_t ce = cudaGetLastError();
    return ce == cudaSuccess ? DTFFT_SUCCESS : cuda_error(ce);
}

int PeerRegistry::reset_barriers() {
    if (!inited_ || !available_ || !flags_ || world_.size() == 1) return DTFFT_SUCCESS;
    cudaError_t ce = cudaDeviceSynchronize();  // my barriers have completed ...
    if (ce != cudaSuccess) return cuda_error(ce);
    world_.barrier();                          // ... and so have everybody's: nobody reads or writes flags now
    for (auto& kv : groups_) {
        if (kv.second.d_peer_flags) cudaFree(kv.second.d_peer_flags);
        if (kv.second.d_members) cudaFree(kv.second.d_members);
        if (kv.second.d_epoch) cudaFree(kv.second.d_epoch);
    }
    groups_.clear();
    ce = cudaMemset(flags_, 0, (size_t)kChannels * world_.size() * sizeof(uint64_t));  // the error word stays
    if (ce != cudaSuccess) return cuda_error(ce);
    ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) return cuda_error(ce);
    world_.barrier();  // no peer may signal before my flags are clean
    return DTFFT_SUCCESS;
}

int PeerRegistry::error_state() const {
    if (!h_err_) return 0;
    return (int)*reinterpret_cast<volatile const uint64_t*>(h_err_);
}

void PeerRegistry::destroy() {
    if (!inited_) return;
    for (auto& kv : groups_) {
        if (kv.second.d_peer_flags) cudaFree(kv.second.d_peer_flags);
        if (kv.second.d_members) cudaFree(kv.second.d_members);
        if (kv.second.d_epoch) cudaFree(kv.second.d_epoch);
    }
    groups_.clear();
    bool any = false;
    for (auto& s : slots_) any |= s.live;
    if (any && world_.size() > 1) {
        cudaDeviceSynchronize();
        world_.barrier();
        for (auto& s : slots_)
            if (s.live)
                for (void* p : s.opened)
                    if (p) ipc_close(p);
        world_.barrier();
    }
    slots_.clear();
    if (flags_) cudaFree(flags_);
    flags_ = nullptr;
    flags_slot_ = -1;
    if (h_err_) cudaFreeHost(h_err_);
    h_err_ = nullptr;
    d_err_host_ = nullptr;
    cudaGetLastError();
    inited_ = false;
    available_ = false;
}

}  // namespace dtfftb
