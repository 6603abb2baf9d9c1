! Reference-side binding for libdtfft_b200.so (iso_c_binding shim; see INTEGRATION.md for where it plugs into dtFFT).
! Not compiled in this repository: the build image has no Fortran compiler.
module dtfft_backend_b200                       ! stands in for src/dtfft_backend_nccl.F90
use iso_c_binding
use dtfft_abstract_backend
  interface
    integer(c_int) function dtfftb_backend_create(backend, backend_type, nccl_comm, comm_rank, comm_size, comm_mapping, &
                                                  send_displs, send_counts, recv_displs, recv_counts, base_storage) bind(C)
      import
      type(c_ptr)                :: backend                       ! dtfftb_backend_t*
      integer(c_int),     value  :: backend_type                  ! dtfft_backend_t%val: 24 NCCL, 27 NCCL_PIPELINED
      type(c_ptr),        value  :: nccl_comm                     ! helper%nccl_comm%member
      integer(c_int),     value  :: comm_rank, comm_size
      integer(c_int32_t)         :: comm_mapping(*)               ! helper%comm_mappings(comm_id)%ranks(0:)
      integer(c_int64_t)         :: send_displs(*), send_counts(*), recv_displs(*), recv_counts(*)  ! elements, 0-based
      integer(c_int64_t), value  :: base_storage
    end function
    integer(c_int) function dtfftb_backend_set_unpack_kernel(backend, kernel) bind(C)
      import;  type(c_ptr), value :: backend, kernel              ! kernel = kernel_b200%handle
    end function
    integer(c_int) function dtfftb_backend_execute(backend, in, out, stream, aux) bind(C)
      import;  type(c_ptr), value :: backend, in, out, stream, aux
    end function
    integer(c_int) function dtfftb_backend_destroy(backend) bind(C)
      import;  type(c_ptr) :: backend
    end function
  end interface
  type, extends(abstract_backend) :: backend_b200
    type(c_ptr) :: handle = c_null_ptr
  contains
    procedure :: create_private  => create     ! (self, helper, base_storage): forwards counts/displs the base class computed
    procedure :: execute_private => execute    ! (self, in, out, stream, aux, error_code): c_loc(in), c_loc(out), c_loc(aux)
    procedure :: destroy_private => destroy
  end type
end module
