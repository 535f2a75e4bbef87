! eigb200_mg -- multi-GPU variants of the drop-in entry points for an MPI caller (one rank per GPU, ONE problem).
! No reference analogue (NVIDIA/Eigensolver_gpu is single-GPU); same dummy-argument lists as zhegvdx_gpu / dsygvdx_gpu
! (lib_eigsolve/zhegvdx_gpu.F90:75-76, dsygvdx_gpu.F90:71-72).  Source only: compile with the caller's nvfortran + MPI.
!
!   call eigb200_mg_setup(MPI_COMM_WORLD)          ! once, after cudaSetDevice(local rank)
!   call zhegvdx_gpu_mg(N, A, lda, B, ldb, Z, ldz, il, iu, w, work, lwork, rwork, lrwork, work_h, lwork_h, &
!                       rwork_h, lrwork_h, iwork_h, liwork_h, Z_h, ldz_h, w_h, info, _skip_host_copy)
!
! Every rank passes the SAME A, B (replicated device inputs); on return every rank holds w(1:N) and Z(:, 1:iu-il+1).
module eigb200_mg
  use cudafor
  use iso_c_binding
  implicit none
contains
  subroutine eigb200_mg_setup(comm)
    use eigb200_c
    use mpi
    implicit none
    integer :: comm, rank, nranks, ierr, istat
    character(kind=c_char), dimension(128) :: id
    call MPI_Comm_rank(comm, rank, ierr)
    call MPI_Comm_size(comm, nranks, ierr)
    if (rank == 0) istat = eigb200_mg_unique_id(id)
    call MPI_Bcast(id, 128, MPI_BYTE, 0, comm, ierr)
    istat = eigb200_mg_init(rank, nranks, id)
    if (istat /= 0) print*, "eigb200_mg_init failed"
  end subroutine eigb200_mg_setup

  subroutine zhegvdx_gpu_mg(N, A, lda, B, ldb, Z, ldz, il, iu, w, work, lwork, rwork, lrwork, &
                            work_h, lwork_h, rwork_h, lrwork_h, iwork_h, liwork_h, Z_h, ldz_h, w_h, info, _skip_host_copy)
    use eigb200_c
    implicit none
    integer                                             :: N, lda, ldb, ldz, il, iu, ldz_h, info
    integer                                             :: lwork_h, lrwork_h, liwork_h, lwork, lrwork, istat, iskip
    real(8), dimension(1:lrwork), device, target        :: rwork
    real(8), dimension(1:lrwork_h), pinned, target      :: rwork_h
    complex(8), dimension(1:lwork), device, target      :: work
    complex(8), dimension(1:lwork_h), pinned, target    :: work_h
    integer, dimension(1:liwork_h), pinned, target      :: iwork_h
    logical, optional                                   :: _skip_host_copy
    complex(8), dimension(1:lda, 1:N), device, target   :: A
    complex(8), dimension(1:ldb, 1:N), device, target   :: B
    complex(8), dimension(1:ldz, 1:N), device, target   :: Z
    complex(8), dimension(1:ldz_h, 1:N), pinned, target :: Z_h
    real(8), dimension(1:N), device, target             :: w
    real(8), dimension(1:N), pinned, target             :: w_h
    iskip = 0
    if (present(_skip_host_copy)) then
      if (_skip_host_copy) iskip = 1
    endif
    istat = eigb200_zhegvdx_mg(N, c_devloc(A), lda, c_devloc(B), ldb, c_devloc(Z), ldz, il, iu, c_devloc(w), &
                               c_devloc(work), lwork, c_devloc(rwork), lrwork, c_loc(work_h), lwork_h, c_loc(rwork_h), &
                               lrwork_h, c_loc(iwork_h), liwork_h, c_loc(Z_h), ldz_h, c_loc(w_h), info, iskip)
  end subroutine zhegvdx_gpu_mg

  subroutine dsygvdx_gpu_mg(N, A, lda, B, ldb, Z, ldz, il, iu, w, work, lwork, &
                            work_h, lwork_h, iwork_h, liwork_h, Z_h, ldz_h, w_h, info, _skip_host_copy)
    use eigb200_c
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
    istat = eigb200_dsygvdx_mg(N, c_devloc(A), lda, c_devloc(B), ldb, c_devloc(Z), ldz, il, iu, c_devloc(w), &
                               c_devloc(work), lwork, c_loc(work_h), lwork_h, c_loc(iwork_h), liwork_h, c_loc(Z_h), &
                               ldz_h, c_loc(w_h), info, iskip)
  end subroutine dsygvdx_gpu_mg
end module eigb200_mg
