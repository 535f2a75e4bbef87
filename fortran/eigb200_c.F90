! eigb200_c -- ISO_C_BINDING interfaces to libeigb200.so (include/eigb200.h).
! Shipped as source (no Fortran compiler in the build image): compile with the caller's nvfortran -cuda next to
! the shim modules of this directory and link with -leigb200.
! Device arrays are passed as type(c_devptr) (cudafor), host arrays as type(c_ptr).
module eigb200_c
  use iso_c_binding
  use cudafor
  implicit none
  interface
    integer(c_int) function eigb200_init() bind(C, name="eigb200_init")
      import :: c_int
    end function eigb200_init

    integer(c_int) function eigb200_dsygvdx(n, A, lda, B, ldb, Z, ldz, il, iu, w, work, lwork, work_h, lwork_h, &
                                            iwork_h, liwork_h, Z_h, ldz_h, w_h, info, skip_host_copy) &
                                            bind(C, name="eigb200_dsygvdx")
      import :: c_int, c_ptr, c_devptr
      integer(c_int), value :: n, lda, ldb, ldz, il, iu, lwork, lwork_h, liwork_h, ldz_h, skip_host_copy
      type(c_devptr), value :: A, B, Z, w, work
      type(c_ptr), value    :: work_h, iwork_h, Z_h, w_h
      integer(c_int)        :: info
    end function eigb200_dsygvdx

    integer(c_int) function eigb200_zhegvdx(n, A, lda, B, ldb, Z, ldz, il, iu, w, work, lwork, rwork, lrwork, &
                                            work_h, lwork_h, rwork_h, lrwork_h, iwork_h, liwork_h, Z_h, ldz_h, w_h, &
                                            info, skip_host_copy) bind(C, name="eigb200_zhegvdx")
      import :: c_int, c_ptr, c_devptr
      integer(c_int), value :: n, lda, ldb, ldz, il, iu, lwork, lrwork, lwork_h, lrwork_h, liwork_h, ldz_h
      integer(c_int), value :: skip_host_copy
      type(c_devptr), value :: A, B, Z, w, work, rwork
      type(c_ptr), value    :: work_h, rwork_h, iwork_h, Z_h, w_h
      integer(c_int)        :: info
    end function eigb200_zhegvdx

    integer(c_int) function eigb200_dsyevd(il, iu, n, A, lda, Z, ldz, w, work, lwork, work_h, lwork_h, iwork_h, &
                                           liwork_h, Z_h, ldz_h, w_h, info) bind(C, name="eigb200_dsyevd")
      import :: c_int, c_ptr, c_devptr
      integer(c_int), value :: il, iu, n, lda, ldz, lwork, lwork_h, liwork_h, ldz_h
      type(c_devptr), value :: A, Z, w, work
      type(c_ptr), value    :: work_h, iwork_h, Z_h, w_h
      integer(c_int)        :: info
    end function eigb200_dsyevd

    integer(c_int) function eigb200_zheevd(il, iu, n, A, lda, Z, ldz, w, work, lwork, rwork, lrwork, work_h, lwork_h, &
                                           rwork_h, lrwork_h, iwork_h, liwork_h, Z_h, ldz_h, w_h, info) &
                                           bind(C, name="eigb200_zheevd")
      import :: c_int, c_ptr, c_devptr
      integer(c_int), value :: il, iu, n, lda, ldz, lwork, lrwork, lwork_h, lrwork_h, liwork_h, ldz_h
      type(c_devptr), value :: A, Z, w, work, rwork
      type(c_ptr), value    :: work_h, rwork_h, iwork_h, Z_h, w_h
      integer(c_int)        :: info
    end function eigb200_zheevd

    ! ---- multi-GPU (one MPI rank per GPU, ONE problem): include/eigb200.h, "library-owned communicator" ----------------
    ! rank 0: eigb200_mg_unique_id(id); MPI_Bcast(id, 128, MPI_BYTE, 0, comm); every rank: eigb200_mg_init(rank, nranks, id)
    integer(c_int) function eigb200_mg_unique_id(id128) bind(C, name="eigb200_mg_unique_id")
      import :: c_int, c_char
      character(kind=c_char), dimension(128) :: id128
    end function eigb200_mg_unique_id

    integer(c_int) function eigb200_mg_init(rank, world, id128) bind(C, name="eigb200_mg_init")
      import :: c_int, c_char
      integer(c_int), value :: rank, world
      character(kind=c_char), dimension(128) :: id128
    end function eigb200_mg_init

    integer(c_int) function eigb200_mg_finalize() bind(C, name="eigb200_mg_finalize")
      import :: c_int
    end function eigb200_mg_finalize

    integer(c_int) function eigb200_dsygvdx_mg(n, A, lda, B, ldb, Z, ldz, il, iu, w, work, lwork, work_h, lwork_h, &
                                               iwork_h, liwork_h, Z_h, ldz_h, w_h, info, skip_host_copy) &
                                               bind(C, name="eigb200_dsygvdx_mg")
      import :: c_int, c_ptr, c_devptr
      integer(c_int), value :: n, lda, ldb, ldz, il, iu, lwork, lwork_h, liwork_h, ldz_h, skip_host_copy
      type(c_devptr), value :: A, B, Z, w, work
      type(c_ptr), value    :: work_h, iwork_h, Z_h, w_h
      integer(c_int)        :: info
    end function eigb200_dsygvdx_mg

    integer(c_int) function eigb200_zhegvdx_mg(n, A, lda, B, ldb, Z, ldz, il, iu, w, work, lwork, rwork, lrwork, &
                                               work_h, lwork_h, rwork_h, lrwork_h, iwork_h, liwork_h, Z_h, ldz_h, w_h, &
                                               info, skip_host_copy) bind(C, name="eigb200_zhegvdx_mg")
      import :: c_int, c_ptr, c_devptr
      integer(c_int), value :: n, lda, ldb, ldz, il, iu, lwork, lrwork, lwork_h, lrwork_h, liwork_h, ldz_h
      integer(c_int), value :: skip_host_copy
      type(c_devptr), value :: A, B, Z, w, work, rwork
      type(c_ptr), value    :: work_h, rwork_h, iwork_h, Z_h, w_h
      integer(c_int)        :: info
    end function eigb200_zhegvdx_mg

    ! optional: stream the solver issues its work on / one-shot "A has been uploaded" event (include/eigb200.h)
    integer(c_int) function eigb200_set_stream(stream) bind(C, name="eigb200_set_stream")
      import :: c_int, cuda_stream_kind
      integer(kind=cuda_stream_kind), value :: stream
    end function eigb200_set_stream

    integer(c_int) function eigb200_set_a_ready_event(ev) bind(C, name="eigb200_set_a_ready_event")
      import :: c_int, cudaEvent
      type(cudaEvent), value :: ev
    end function eigb200_set_a_ready_event

    integer(c_int) function eigb200_set_option(name, val) bind(C, name="eigb200_set_option")
      import :: c_int, c_char
      character(kind=c_char), dimension(*) :: name      ! null-terminated, e.g. "nvtx"//c_null_char
      integer(c_int), value :: val
    end function eigb200_set_option
  end interface
end module eigb200_c
