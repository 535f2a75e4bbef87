! Drop-in for module dsygvdx_gpu (reference: lib_eigsolve/dsygvdx_gpu.F90:24-170).
module dsygvdx_gpu
  use cudafor
  use iso_c_binding
  implicit none
contains
  subroutine dsygvdx_gpu(N, A, lda, B, ldb, Z, ldz, il, iu, w, work, lwork, &
                         work_h, lwork_h, iwork_h, liwork_h, Z_h, ldz_h, w_h, info, _skip_host_copy)
    use eigb200_c
    use eigsolve_vars
    implicit none
    integer                                          :: N, lda, ldb, ldz, il, iu, ldz_h, info
    integer                                          :: lwork_h, liwork_h, lwork, istat, iskip
    real(8), dimension(1:lwork), device, target      :: work
    real(8), dimension(1:lwork_h), pinned, target    :: work_h
    integer, dimension(1:liwork_h), pinned, target   :: iwork_h
    logical, optional                                :: _skip_host_copy
    real(8), dimension(1:lda, 1:N), device, target   :: A
    real(8), dimension(1:ldb, 1:N), device, target   :: B
    real(8), dimension(1:ldz, 1:N), device, target   :: Z
    real(8), dimension(1:ldz_h, 1:N), pinned, target :: Z_h
    real(8), dimension(1:N), device, target          :: w
    real(8), dimension(1:N), pinned, target          :: w_h

    iskip = 0
    if (present(_skip_host_copy)) then
      if (_skip_host_copy) iskip = 1
    endif
    if (initialized == 0) call init_eigsolve_gpu
    istat = eigb200_dsygvdx(N, c_devloc(A), lda, c_devloc(B), ldb, c_devloc(Z), ldz, il, iu, c_devloc(w), &
                            c_devloc(work), lwork, c_loc(work_h), lwork_h, c_loc(iwork_h), liwork_h, c_loc(Z_h), ldz_h, &
                            c_loc(w_h), info, iskip)
  end subroutine dsygvdx_gpu
end module dsygvdx_gpu
