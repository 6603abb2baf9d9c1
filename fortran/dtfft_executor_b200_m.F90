! Reference-side binding for libdtfft_b200.so (iso_c_binding shim; see INTEGRATION.md for where it plugs into dtFFT).
! Not compiled in this repository: the build image has no Fortran compiler.
module dtfft_executor_b200_m                    ! stands in for src/interfaces/fft/cufft/dtfft_executor_cufft_m.F90
use iso_c_binding
use dtfft_abstract_executor
  interface
    integer(c_int) function dtfftb_executor_create(executor, fft_rank, fft_type, precision, idist, odist, how_many, &
                                                   fft_sizes, inembed, onembed, stream) bind(C)
      import
      type(c_ptr)               :: executor
      integer(c_int),     value :: fft_rank, fft_type, precision  ! FFT_C2C = 0, FFT_R2C = 1; dtfft_precision_t%val
      integer(c_int32_t), value :: idist, odist, how_many
      integer(c_int32_t)        :: fft_sizes(*), inembed(*), onembed(*)
      type(c_ptr),        value :: stream                          ! get_conf_stream()
    end function
    integer(c_int) function dtfftb_executor_execute(executor, a, b, sign) bind(C)
      import;  type(c_ptr), value :: executor, a, b;  integer(c_int), value :: sign   ! FFT_FORWARD = -1, FFT_BACKWARD = +1
    end function
    integer(c_int) function dtfftb_executor_destroy(executor) bind(C)
      import;  type(c_ptr) :: executor
    end function
  end interface
  type, extends(abstract_executor) :: b200_executor
    type(c_ptr) :: handle = c_null_ptr
  contains
    procedure :: create_private  => create     ! same argument list as create_interface, :67-84
    procedure :: execute_private => execute
    procedure :: destroy_private => destroy
    procedure, nopass :: mem_alloc, mem_free   ! cudaMalloc / cudaFree as today
  end type
end module
