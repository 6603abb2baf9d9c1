! Reference-side binding for libdtfft_b200.so (iso_c_binding shim; see INTEGRATION.md, level 2).
! Not compiled in this repository: the build image has no Fortran compiler.
!
! Interfaces of the plan-level C ABI (include/dtfft_b200_api.h) for a Fortran caller that replaces the
! whole CUDA-platform plan: the names, argument order and enum values are those of the reference's C API
! (include/dtfft.h:397-1407), so the bodies of dtfft_plan_c2c_t%create / execute / get_local_sizes /
! mem_alloc (src/dtfft_plan.F90) reduce to one call each.  `comm` is a dtfftb_comm_t* built by a small C
! helper from MPI_Comm_f2c(comm%MPI_VAL) (include/dtfft_b200_mpi.h: dtfftb_comm_from_mpi).
module dtfft_b200_api
use iso_c_binding
implicit none
public

  interface
    integer(c_int32_t) function dtfft_create_plan_c2c(ndims, dims, comm, precision, effort, executor, plan) bind(C)
      import
      integer(c_int8_t),  value :: ndims
      integer(c_int32_t)        :: dims(*)
      type(c_ptr),        value :: comm
      integer(c_int),     value :: precision, effort, executor
      type(c_ptr)               :: plan
    end function
    integer(c_int32_t) function dtfft_create_plan_r2c(ndims, dims, comm, precision, effort, executor, plan) bind(C)
      import
      integer(c_int8_t),  value :: ndims
      integer(c_int32_t)        :: dims(*)
      type(c_ptr),        value :: comm
      integer(c_int),     value :: precision, effort, executor
      type(c_ptr)               :: plan
    end function
    integer(c_int32_t) function dtfft_create_plan_r2r(ndims, dims, kinds, comm, precision, effort, executor, plan) bind(C)
      import
      integer(c_int8_t),  value :: ndims
      integer(c_int32_t)        :: dims(*)
      integer(c_int)            :: kinds(*)
      type(c_ptr),        value :: comm
      integer(c_int),     value :: precision, effort, executor
      type(c_ptr)               :: plan
    end function
    integer(c_int32_t) function dtfft_execute(plan, in, out, execute_type, aux) bind(C)
      import
      type(c_ptr), value    :: plan, in, out, aux
      integer(c_int), value :: execute_type
    end function
    integer(c_int32_t) function dtfft_transpose(plan, in, out, transpose_type, aux) bind(C)
      import
      type(c_ptr), value    :: plan, in, out, aux
      integer(c_int), value :: transpose_type
    end function
    integer(c_int32_t) function dtfft_reshape(plan, in, out, reshape_type, aux) bind(C)
      import
      type(c_ptr), value    :: plan, in, out, aux
      integer(c_int), value :: reshape_type
    end function
    integer(c_int32_t) function dtfft_destroy(plan) bind(C)
      import
      type(c_ptr) :: plan            ! dtfft_plan_t*: set to NULL on return
    end function
    integer(c_int32_t) function dtfft_get_local_sizes(plan, in_starts, in_counts, out_starts, out_counts, alloc_size) bind(C)
      import
      type(c_ptr), value :: plan
      integer(c_int32_t) :: in_starts(*), in_counts(*), out_starts(*), out_counts(*)
      integer(c_size_t)  :: alloc_size
    end function
    integer(c_int32_t) function dtfft_get_alloc_bytes(plan, alloc_bytes) bind(C)
      import
      type(c_ptr), value :: plan
      integer(c_size_t)  :: alloc_bytes
    end function
    integer(c_int32_t) function dtfft_get_aux_bytes(plan, aux_bytes) bind(C)
      import
      type(c_ptr), value :: plan
      integer(c_size_t)  :: aux_bytes
    end function
    integer(c_int32_t) function dtfft_mem_alloc(plan, alloc_bytes, ptr) bind(C)
      import
      type(c_ptr), value       :: plan
      integer(c_size_t), value :: alloc_bytes
      type(c_ptr)              :: ptr
    end function
    integer(c_int32_t) function dtfft_mem_free(plan, ptr) bind(C)
      import
      type(c_ptr), value :: plan, ptr
    end function
    integer(c_int32_t) function dtfft_get_stream(plan, stream) bind(C)
      import
      type(c_ptr), value :: plan
      type(c_ptr)        :: stream   ! cudaStream_t
    end function
  end interface
end module dtfft_b200_api
