/*
 * dtfft_b200_mpi.h -- header-only adapter for MPI programs (the build image has no MPI: the header
 * is compiled and exercised against a single-process stand-in, tests/c/mpi_stub/mpi.h).  Turns an MPI_Comm into the dtfftb_comm_t the library takes in
 * place of the reference's `MPI_Comm comm` argument (include/dtfft.h:397-515 of the reference):
 *
 *     dtfftb_mpi_comm_t c;
 *     dtfftb_comm_from_mpi(MPI_COMM_WORLD, &c);
 *     dtfft_create_plan_c2c(3, dims, &c.comm, DTFFT_DOUBLE, DTFFT_ESTIMATE, DTFFT_EXECUTOR_NONE, &plan);
 *
 * The plan keeps `c.comm.ctx` (= &c) for its later collectives (dtfft_mem_alloc with the NVLINK_FUSED
 * backend, dtfft_destroy): `c` must outlive every plan created from it, like the MPI_Comm itself.
 *
 * A cartesian communicator (MPI_Cart_create) is forwarded as cart_ndims / cart_dims in dtFFT's
 * order (dims[0] fastest = LAST dimension of the MPI grid is NOT reversed: the reference reads
 * MPI_Cart_get verbatim, src/dtfft_transpose_plan.F90:139-158).
 */
#ifndef DTFFT_B200_MPI_H
#define DTFFT_B200_MPI_H

#include <mpi.h>

#include "dtfft_b200_api.h"

typedef struct {
    dtfftb_comm_t comm;
    MPI_Comm mpi;
} dtfftb_mpi_comm_t;

static int dtfftb_mpi_allgather_(void* ctx, const void* send, void* recv, int64_t bytes) {
    MPI_Comm c = ((dtfftb_mpi_comm_t*)ctx)->mpi;
    return MPI_Allgather(send, (int)bytes, MPI_BYTE, recv, (int)bytes, MPI_BYTE, c) == MPI_SUCCESS ? 0 : 1;
}

static inline void dtfftb_comm_from_mpi(MPI_Comm mpi, dtfftb_mpi_comm_t* out) {
    int rank = 0, size = 1, topo = MPI_UNDEFINED;
    MPI_Comm_rank(mpi, &rank);
    MPI_Comm_size(mpi, &size);
    out->mpi = mpi;
    out->comm.rank = rank;
    out->comm.size = size;
    out->comm.ctx = out;
    out->comm.allgather = dtfftb_mpi_allgather_;
    out->comm.cart_ndims = 0;
    out->comm.cart_dims[0] = out->comm.cart_dims[1] = out->comm.cart_dims[2] = 1;
    MPI_Topo_test(mpi, &topo);
    if (topo == MPI_CART) {
        int nd = 0, dims[3] = {1, 1, 1}, periods[3], coords[3];
        MPI_Cartdim_get(mpi, &nd);
        if (nd >= 1 && nd <= 3) {
            MPI_Cart_get(mpi, nd, dims, periods, coords);
            out->comm.cart_ndims = nd;
            for (int i = 0; i < nd; ++i) out->comm.cart_dims[i] = dims[i];
        }
    }
}

#endif /* DTFFT_B200_MPI_H */
