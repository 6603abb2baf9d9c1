// Process-group plumbing of the plan layer.
//
// The reference bootstraps everything through MPI (communicator splitting, Allgather of
// pencil extents, broadcast of the ncclUniqueId: src/dtfft_abstract_backend.F90:395-457,
// src/dtfft_reshape_handle_generic.F90:143-144).  Here the host program hands the library a
// tiny vtable (dtfftb_comm_t, include/dtfft_b200.h) with ONE collective -- allgather of a
// fixed number of bytes per rank -- and every piece of metadata exchange is expressed with
// it.  An MPI program wraps MPI_Allgather; the Python host wraps torch.distributed
// (gloo on CPU, NCCL on GPUs); a NULL comm is a single-rank world.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/dtfft_b200.h"

namespace dtfftb {

class Comm {
public:
    Comm() = default;
    explicit Comm(const dtfftb_comm_t* c) {
        if (c && c->size > 1) {
            c_ = *c;
            has_ = true;
        }
    }
    const dtfftb_comm_t* raw() const { return has_ ? &c_ : nullptr; }
    int rank() const { return has_ ? c_.rank : 0; }
    int size() const { return has_ ? c_.size : 1; }
    // recv must hold size() * bytes.  Returns 0 or a non-zero host error.
    int allgather(const void* send, void* recv, int64_t bytes) const {
        if (!has_) {
            std::memcpy(recv, send, (size_t)bytes);
            return 0;
        }
        return c_.allgather(c_.ctx, send, recv, bytes);
    }
    template <typename T>
    int allgather_v(const T& mine, std::vector<T>& all) const {
        all.resize((size_t)size());
        return allgather(&mine, all.data(), (int64_t)sizeof(T));
    }
    int barrier() const {
        char c = 0;
        std::vector<char> r((size_t)size());
        return allgather(&c, r.data(), 1);
    }
    // max over ranks (used like the reference's MPI_Allreduce(MAX) on error codes / timings)
    double max(double v) const {
        std::vector<double> all;
        allgather_v(v, all);
        double m = all[0];
        for (double x : all) m = x > m ? x : m;
        return m;
    }
    long long sum(long long v) const {
        std::vector<long long> all;
        allgather_v(v, all);
        long long s = 0;
        for (long long x : all) s += x;
        return s;
    }

private:
    dtfftb_comm_t c_{};
    bool has_ = false;
};

}  // namespace dtfftb
