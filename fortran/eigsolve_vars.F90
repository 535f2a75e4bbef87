! Drop-in for module eigsolve_vars (reference: lib_eigsolve/eigsolve_vars.F90:25-61).
! The reference keeps cuBLAS/cuSOLVER handles, streams and events here; eigb200 owns its context inside
! libeigb200.so, so only `initialized` and `init_eigsolve_gpu` remain for callers that use them
! (test_driver/test_zhegvdx.F90:79,141; zhegvdx_gpu.F90:131).
module eigsolve_vars
  use eigb200_c
  implicit none
  integer :: initialized = 0
contains
  subroutine init_eigsolve_gpu()
    integer :: istat
    istat = eigb200_init()
    if (istat == 0) initialized = 1
  end subroutine init_eigsolve_gpu
end module eigsolve_vars
