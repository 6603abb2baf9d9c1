! Reference-side binding for libdtfft_b200.so (iso_c_binding shim; see INTEGRATION.md for where it plugs into dtFFT).
! Not compiled in this repository: the build image has no Fortran compiler.
module dtfft_kernel_b200
use iso_c_binding
use iso_fortran_env
use dtfft_abstract_kernel
use dtfft_parameters
use dtfft_interface_cuda_runtime, only: dtfft_stream_t
implicit none
private
public :: kernel_b200

  interface
    integer(c_int) function dtfftb_kernel_create(kernel, ndims, dims, kernel_type, base_storage, &
                                                 neighbor_data, n_neighbors, effort, force_effort) bind(C)
      import
      type(c_ptr)                :: kernel          ! dtfftb_kernel_t*
      integer(c_int),     value  :: ndims
      integer(c_int32_t)         :: dims(*)
      integer(c_int),     value  :: kernel_type     ! kernel_type_t%val (same numbering)
      integer(c_int64_t), value  :: base_storage
      type(c_ptr),        value  :: neighbor_data   ! c_loc(neighbor_data(1,1)) : (5, P) column-major, or c_null_ptr
      integer(c_int),     value  :: n_neighbors, effort, force_effort
    end function
    integer(c_int) function dtfftb_kernel_execute(kernel, in, out, stream, neighbor, sync) bind(C)
      import
      type(c_ptr), value    :: kernel, in, out, stream
      integer(c_int), value :: neighbor            ! 1-based, 0 = not given
      integer(c_int), value :: sync
    end function
    integer(c_int) function dtfftb_kernel_destroy(kernel) bind(C)
      import
      type(c_ptr) :: kernel
    end function
    ! One launch for every neighbour (what the non-pipelined unpack of reshape_handle_generic wants instead of the
    ! loop of P launches, src/dtfft_kernel_device.F90:167-174).
    integer(c_int) function dtfftb_kernel_execute_all(kernel, in, out, stream) bind(C)
      import
      type(c_ptr), value :: kernel, in, out, stream
    end function
    ! Kernel over explicit boxes (10 x n int64: n0 n1 n2 in_off out_off is1 is2 os0 os1 os2, elements) with one
    ! destination base per box: the direct-store transposition of a fused backend (the reference's fused backends
    ! drive pack_forward / pack_backward per peer, src/dtfft_reshape_handle_generic.F90:447).
    integer(c_int) function dtfftb_kernel_create_boxes(kernel, family, base_storage, n_boxes, boxes, out_bases) bind(C)
      import
      type(c_ptr)                :: kernel
      integer(c_int),     value  :: family         ! 2 = tiled transpose, 3 = row copy
      integer(c_int64_t), value  :: base_storage
      integer(c_int),     value  :: n_boxes
      integer(c_int64_t)         :: boxes(10, *)
      type(c_ptr),        value  :: out_bases       ! c_loc of n_boxes c_ptr (peer-mapped bases), or c_null_ptr
    end function
    ! Timed tile autotune with the per-candidate report of src/dtfft_kernel_device.F90:385-389 (ms and GB/s).
    integer(c_int) function dtfftb_kernel_autotune_report(kernel, in, out, stream, n_warmup, n_iters, max_entries, &
                                                          n_entries, tiles, ms, gbs) bind(C)
      import
      type(c_ptr),    value :: kernel, in, out, stream
      integer(c_int), value :: n_warmup, n_iters, max_entries
      integer(c_int)        :: n_entries
      integer(c_int32_t)    :: tiles(3, *)
      real(c_float)         :: ms(*)
      real(c_double)        :: gbs(*)
    end function
  end interface

  type, extends(abstract_kernel) :: kernel_b200
    type(c_ptr) :: handle = c_null_ptr
  contains
    procedure :: create_private  => create
    procedure :: execute_private => execute
    procedure :: destroy_private => destroy
  end type

contains
  subroutine create(self, effort, base_storage, force_effort)          ! replaces kernel_device%create, :61-102
    class(kernel_b200), intent(inout) :: self
    type(dtfft_effort_t), intent(in)  :: effort
    integer(int64),       intent(in)  :: base_storage
    logical, optional,    intent(in)  :: force_effort
    integer(c_int) :: ierr, nn, fe
    type(c_ptr) :: nd
    nd = c_null_ptr; nn = 0; fe = 0
    if ( allocated(self%neighbor_data) ) then
      nd = c_loc(self%neighbor_data); nn = size(self%neighbor_data, 2)
    endif
    if ( present(force_effort) ) then; if ( force_effort ) fe = 1; endif
    ierr = dtfftb_kernel_create(self%handle, size(self%dims), self%dims, self%kernel_type%val, base_storage, nd, nn, effort%val, fe)
    if ( ierr /= 0 ) INTERNAL_ERROR("dtfftb_kernel_create failed")
  end subroutine
  subroutine execute(self, in, out, stream, sync, neighbor)            ! replaces kernel_device%execute, :104-178
    class(kernel_b200), intent(inout) :: self
    type(c_ptr),          intent(in)  :: in, out
    type(dtfft_stream_t), intent(in)  :: stream
    logical,              intent(in)  :: sync
    integer(int32), optional, intent(in) :: neighbor
    integer(c_int) :: ierr, nb, sy
    nb = 0; if ( present(neighbor) ) nb = neighbor
    sy = 0; if ( sync ) sy = 1
    ierr = dtfftb_kernel_execute(self%handle, in, out, stream%stream, nb, sy)
    if ( ierr /= 0 ) INTERNAL_ERROR("dtfftb_kernel_execute failed")
  end subroutine
  subroutine destroy(self)
    class(kernel_b200), intent(inout) :: self
    integer(c_int) :: ierr
    ierr = dtfftb_kernel_destroy(self%handle)
  end subroutine
end module
