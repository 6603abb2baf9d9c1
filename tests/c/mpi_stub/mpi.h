/* Single-process stand-in for <mpi.h>, just enough to compile and run include/dtfft_b200_mpi.h in a
 * test (the image has no MPI).  Communicator 0 = MPI_COMM_WORLD (one rank, no topology);
 * communicator 1 = a 1 x 1 x 1 cartesian communicator.  Test infrastructure only. */
#ifndef DTFFTB_TEST_MPI_STUB_H
#define DTFFTB_TEST_MPI_STUB_H
#include <string.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_BYTE 1
#define MPI_SUCCESS 0
#define MPI_UNDEFINED (-32766)
#define MPI_CART 2

static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return MPI_SUCCESS; }
static inline int MPI_Allgather(const void* s, int n, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
    (void)st; (void)rt; (void)c;
    if (n != rn) return 1;
    memcpy(r, s, (size_t)n);
    return MPI_SUCCESS;
}
static inline int MPI_Topo_test(MPI_Comm c, int* t) { *t = c == 1 ? MPI_CART : MPI_UNDEFINED; return MPI_SUCCESS; }
static inline int MPI_Cartdim_get(MPI_Comm c, int* nd) { *nd = c == 1 ? 3 : 0; return MPI_SUCCESS; }
static inline int MPI_Cart_get(MPI_Comm c, int nd, int* dims, int* periods, int* coords) {
    (void)c;
    for (int i = 0; i < nd; ++i) dims[i] = 1, periods[i] = 0, coords[i] = 0;
    return MPI_SUCCESS;
}
#endif
