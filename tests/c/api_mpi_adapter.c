/* include/dtfft_b200_mpi.h compiled against a single-process MPI stand-in (tests/c/mpi_stub/mpi.h):
 * an MPI_Comm becomes the dtfftb_comm_t the plan constructors take, the allgather callback round-trips,
 * a cartesian communicator is forwarded as a process grid.  Host only (dry plans). */
#include <stdio.h>
#include <stdlib.h>

#include "dtfft_b200_mpi.h"

#define EXPECT(c)                                                            \
    do {                                                                     \
        if (!(c)) {                                                          \
            fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);   \
            exit(1);                                                         \
        }                                                                    \
    } while (0)

int main(void) {
    dtfftb_mpi_comm_t c;
    dtfftb_comm_from_mpi(MPI_COMM_WORLD, &c);
    EXPECT(c.comm.rank == 0 && c.comm.size == 1 && c.comm.cart_ndims == 0 && c.comm.ctx == &c);
    long long in = 0x1122334455667788ll, out = 0;
    EXPECT(c.comm.allgather(c.comm.ctx, &in, &out, (int64_t)sizeof in) == 0 && out == in);

    const int32_t dims[3] = {32, 24, 16};
    dtfft_plan_t plan = NULL;
    EXPECT(dtfftb_plan_create_dry(0, 3, dims, NULL, &c.comm, DTFFT_DOUBLE, DTFFT_EXECUTOR_NONE, &plan) == DTFFT_SUCCESS);
    int8_t nd = 0;
    const int32_t* g = NULL;
    EXPECT(dtfft_get_grid_dims(plan, &nd, &g) == DTFFT_SUCCESS && nd == 3 && g[0] == 1 && g[1] == 1 && g[2] == 1);
    EXPECT(dtfft_destroy(&plan) == DTFFT_SUCCESS);

    dtfftb_mpi_comm_t cart;
    dtfftb_comm_from_mpi((MPI_Comm)1, &cart);  /* the stub's 1 x 1 x 1 cartesian communicator */
    EXPECT(cart.comm.cart_ndims == 3 && cart.comm.cart_dims[0] == 1 && cart.comm.cart_dims[1] == 1 && cart.comm.cart_dims[2] == 1);
    EXPECT(dtfftb_plan_create_dry(0, 3, dims, NULL, &cart.comm, DTFFT_DOUBLE, DTFFT_EXECUTOR_NONE, &plan) == DTFFT_SUCCESS);
    EXPECT(dtfft_destroy(&plan) == DTFFT_SUCCESS);
    printf("api_mpi_adapter OK\n");
    return 0;
}
