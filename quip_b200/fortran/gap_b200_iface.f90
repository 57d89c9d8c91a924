! gap_b200_iface.f90 -- ISO_C_BINDING interface to libgapb200.so (include/gap_b200.h) and the replacement bodies
! of IPModel_GAP_Initialise_str / IPModel_GAP_Calc / IPModel_GAP_Finalise a QUIP maintainer would drop into
! src/Potentials/IPModel_GAP.f95 (:149, :233, :194).
!
! NOT COMPILED IN THIS REPOSITORY: the build image has no Fortran compiler.  The C ABI it binds is exercised by
! quip_b200/potential.py (ctypes) and tests/; the array layouts below are the ones those tests use
! (pos(3,N), f(3,N), virial(3,3) column-major, local_virial(9,N)), so no copies or transposes are needed.
module gap_b200_iface
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: gap_potential_initialise, gap_potential_filename_initialise, gap_potential_finalise, gap_potential_cutoff
  public :: gap_potential_calc, gap_potential_set_partition, gap_last_error, gap_b200_error_string
  public :: gap_potential_set_atom_mask, gap_potential_get_energy_per_coordinate, gap_potential_get_local_gap_variance
  public :: gap_potential_set_timing, gap_potential_last_timings, gap_md_run
  public :: gap_comm_get_unique_id, gap_potential_set_comm, gap_potential_set_cutoff_skin, gap_potential_set_resid, GAP_COMM_ID_BYTES

  integer, parameter :: GAP_COMM_ID_BYTES = 128

  interface
     ! int gap_potential_initialise(gap_potential** pot, const char* args_str, const char* param_str, const char* base_dir, int device)
     function gap_potential_initialise(pot, args_str, param_str, base_dir, device) bind(C, name="gap_potential_initialise") result(ierr)
       import :: c_ptr, c_char, c_int
       type(c_ptr), intent(out) :: pot
       character(kind=c_char), dimension(*), intent(in) :: args_str, param_str, base_dir
       integer(c_int), value :: device
       integer(c_int) :: ierr
     end function
     function gap_potential_filename_initialise(pot, args_str, param_filename, device) &
          bind(C, name="gap_potential_filename_initialise") result(ierr)
       import :: c_ptr, c_char, c_int
       type(c_ptr), intent(out) :: pot
       character(kind=c_char), dimension(*), intent(in) :: args_str, param_filename
       integer(c_int), value :: device
       integer(c_int) :: ierr
     end function
     subroutine gap_potential_finalise(pot) bind(C, name="gap_potential_finalise")
       import :: c_ptr
       type(c_ptr), value :: pot
     end subroutine
     function gap_potential_cutoff(pot) bind(C, name="gap_potential_cutoff") result(rc)
       import :: c_ptr, c_double
       type(c_ptr), value :: pot
       real(c_double) :: rc
     end function
     function gap_potential_set_partition(pot, rank, n_ranks) bind(C, name="gap_potential_set_partition") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: pot
       integer(c_int), value :: rank, n_ranks
       integer(c_int) :: ierr
     end function
     ! the reduction over ranks inside the library (replaces sum_in_place, IPModel_GAP.f95:538-556): rank 0 creates the id, the host
     ! broadcasts its 128 bytes (MPI_Bcast), every rank joins; from then on calc returns totals on every rank
     function gap_comm_get_unique_id(id) bind(C, name="gap_comm_get_unique_id") result(ierr)
       import :: c_char, c_int
       character(kind=c_char), intent(out) :: id(128)
       integer(c_int) :: ierr
     end function
     function gap_potential_set_comm(pot, id, rank, n_ranks) bind(C, name="gap_potential_set_comm") result(ierr)
       import :: c_ptr, c_char, c_int
       type(c_ptr), value :: pot
       character(kind=c_char), intent(in) :: id(128)
       integer(c_int), value :: rank, n_ranks
       integer(c_int) :: ierr
     end function
     ! at%cutoff_skin (Connection.f95:1085-1128): neighbour list built to cutoff + skin, reused while max displacement < skin / 2
     function gap_potential_set_cutoff_skin(pot, cutoff_skin) bind(C, name="gap_potential_set_cutoff_skin") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: pot
       real(c_double), value :: cutoff_skin
       integer(c_int) :: ierr
     end function
     ! residue ids for distance_2b only_intra / only_inter (the integer property named by resid_name, descriptors.f95:4660-4668)
     function gap_potential_set_resid(pot, n, resid) bind(C, name="gap_potential_set_resid") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: pot, resid       ! resid: c_loc of integer(c_int) resid(n), or c_null_ptr
       integer(c_int), value :: n
       integer(c_int) :: ierr
     end function
     ! optional inputs / outputs the reference keeps in the Atoms object (IPModel_GAP.f95:324-337, 558-573)
     function gap_potential_set_atom_mask(pot, n, mask) bind(C, name="gap_potential_set_atom_mask") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: pot, mask        ! mask: c_loc of integer(c_int) mask(n), or c_null_ptr to clear it
       integer(c_int), value :: n
       integer(c_int) :: ierr
     end function
     function gap_potential_get_energy_per_coordinate(pot, energy_per_coordinate) &
          bind(C, name="gap_potential_get_energy_per_coordinate") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: pot
       real(c_double), intent(out) :: energy_per_coordinate(*)
       integer(c_int) :: ierr
     end function
     function gap_potential_get_local_gap_variance(pot, n, local_gap_variance, gap_variance_gradient) &
          bind(C, name="gap_potential_get_local_gap_variance") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: pot, gap_variance_gradient   ! c_loc of real(dp) (3,n), or c_null_ptr
       integer(c_int), value :: n
       real(c_double), intent(out) :: local_gap_variance(*)
       integer(c_int) :: ierr
     end function
     ! system_timer analogue: per-stage device milliseconds of the last calc (recorded only after set_timing(pot, 1))
     function gap_potential_set_timing(pot, on) bind(C, name="gap_potential_set_timing") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: pot
       integer(c_int), value :: on
       integer(c_int) :: ierr
     end function
     function gap_potential_last_timings(pot, ms8) bind(C, name="gap_potential_last_timings") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: pot
       real(c_double), intent(out) :: ms8(8)
       integer(c_int) :: ierr
     end function
     ! DynamicalSystem_run (Potential.f95:2304) for plain NVE dynamics, state resident on the GPU for the whole run
     function gap_md_run(pot, n, pos, velo, z, mass, lattice, pbc, dt, n_steps, args_str, epot, ekin) bind(C, name="gap_md_run") result(ierr)
       import :: c_ptr, c_char, c_int, c_double
       type(c_ptr), value :: pot, epot, ekin              ! epot/ekin: c_loc of real(dp)(0:n_steps) or c_null_ptr
       integer(c_int), value :: n, n_steps
       real(c_double), intent(inout) :: pos(3, *), velo(3, *)
       integer(c_int), intent(in) :: z(*), pbc(3)
       real(c_double), intent(in) :: mass(*), lattice(3, 3)
       real(c_double), value :: dt
       character(kind=c_char), dimension(*), intent(in) :: args_str
       integer(c_int) :: ierr
     end function
     ! absent optional outputs are passed as C_NULL_PTR, hence type(c_ptr), value for every output
     function gap_potential_calc(pot, n, pos, z, lattice, pbc, args_str, energy, local_e, force, virial, local_virial) &
          bind(C, name="gap_potential_calc") result(ierr)
       import :: c_ptr, c_char, c_int, c_double
       type(c_ptr), value :: pot
       integer(c_int), value :: n
       real(c_double), intent(in) :: pos(3, *), lattice(3, 3)
       integer(c_int), intent(in) :: z(*), pbc(3)
       character(kind=c_char), dimension(*), intent(in) :: args_str
       type(c_ptr), value :: energy, local_e, force, virial, local_virial
       integer(c_int) :: ierr
     end function
     function gap_last_error() bind(C, name="gap_last_error") result(msg)
       import :: c_ptr
       type(c_ptr) :: msg
     end function
  end interface

contains

  function gap_b200_error_string() result(s)
    character(len=:), allocatable :: s
    character(kind=c_char), pointer :: p(:)
    type(c_ptr) :: cp
    integer :: n
    cp = gap_last_error()
    s = ""
    if (.not. c_associated(cp)) return
    call c_f_pointer(cp, p, [4096])
    n = 0
    do while (n < 4096)
       if (p(n + 1) == c_null_char) exit
       n = n + 1
    end do
    allocate(character(len=n) :: s)
    s = transfer(p(1:n), s)
  end function

end module gap_b200_iface

! ----------------------------------------------------------------------------------------------------------------
! Replacement bodies inside module IPModel_GAP_module (src/Potentials/IPModel_GAP.f95).  type(IPModel_GAP) gains one
! component:   type(c_ptr) :: b200 = c_null_ptr
! ----------------------------------------------------------------------------------------------------------------
!
! subroutine IPModel_GAP_Initialise_str(this, args_str, param_str)          ! IPModel_GAP.f95:149
!   use gap_b200_iface
!   type(IPModel_GAP), intent(inout) :: this
!   character(len=*), intent(in) :: args_str, param_str
!   call Finalise(this)
!   ! Potential_Filename_Initialise has already chdir'ed to the XML's directory (Potential.f95:455-466): base_dir = "."
!   if (gap_potential_initialise(this%b200, trim(args_str)//c_null_char, trim(param_str)//c_null_char, "."//c_null_char, &
!                                gap_b200_device()) /= 0) &
!      call system_abort("IPModel_GAP_Initialise_str: "//gap_b200_error_string())
!   this%cutoff = gap_potential_cutoff(this%b200)                             ! read by IP_cutoff, IP.f95:704-705
!   this%initialised = .true.
! end subroutine
!
! subroutine IPModel_GAP_Finalise(this)                                       ! IPModel_GAP.f95:194
!   if (c_associated(this%b200)) call gap_potential_finalise(this%b200)
!   this%b200 = c_null_ptr ; this%cutoff = 0.0_dp ; this%initialised = .false.
! end subroutine
!
! subroutine IPModel_GAP_Calc(this, at, e, local_e, f, virial, local_virial, args_str, mpi, error)   ! IPModel_GAP.f95:233
!   use gap_b200_iface
!   type(IPModel_GAP), intent(inout) :: this
!   type(Atoms), intent(inout) :: at
!   real(dp), intent(out), optional, target :: e, local_e(:), f(:,:), local_virial(:,:), virial(3,3)
!   character(len=*), intent(in), optional :: args_str
!   type(MPI_Context), intent(in), optional :: mpi
!   integer, intent(out), optional :: error
!   type(c_ptr) :: pe, ple, pf, pv, plv
!   integer(c_int) :: pbc(3), ierr
!   INIT_ERROR(error)
!   pe = c_null_ptr; ple = c_null_ptr; pf = c_null_ptr; pv = c_null_ptr; plv = c_null_ptr
!   if (present(e)) pe = c_loc(e)
!   if (present(local_e)) then; call check_size('Local_E', local_e, (/at%N/), 'IPModel_GAP_Calc', error); ple = c_loc(local_e); end if
!   if (present(f)) then; call check_size('Force', f, (/3, at%N/), 'IPModel_GAP_Calc', error); pf = c_loc(f); end if
!   if (present(virial)) pv = c_loc(virial)
!   if (present(local_virial)) then
!      call check_size('Local_virial', local_virial, (/9, at%N/), 'IPModel_GAP_Calc', error); plv = c_loc(local_virial)
!   end if
!   pbc = merge(1, 0, at%is_periodic)
!   if (present(mpi)) then                                    ! the reference's atom mask (descriptors.f95:1036-1051) + its reduction
!      if (mpi%active .and. .not. this%b200_comm_set) then    ! once per potential: NCCL communicator over the ranks' GPUs
!         if (mpi%my_proc == 0) ierr = gap_comm_get_unique_id(comm_id)
!         call bcast(mpi, comm_id)                            ! MPI_Bcast of the 128 bytes (MPI_context.f95)
!         ierr = gap_potential_set_comm(this%b200, comm_id, mpi%my_proc, mpi%n_procs)
!         this%b200_comm_set = .true.
!      end if
!   end if
!   ierr = gap_potential_set_cutoff_skin(this%b200, at%cutoff_skin)
!   if (has_atom_mask_name) then                              ! :344-346: the logical property becomes an integer mask
!      imask = merge(1_c_int, 0_c_int, atom_mask_pointer)
!      ierr = gap_potential_set_atom_mask(this%b200, at%N, c_loc(imask))
!   end if
!   ierr = gap_potential_calc(this%b200, at%N, at%pos, at%Z, at%lattice, pbc, trim(args_str)//c_null_char, pe, ple, pf, pv, plv)
!   if (ierr /= 0) then
!      RAISE_ERROR("IPModel_GAP_Calc: "//gap_b200_error_string(), error)
!   end if
!   if (do_energy_per_coordinate) then                        ! :573 (sum_in_place over mpi first, :549)
!      ierr = gap_potential_get_energy_per_coordinate(this%b200, energy_per_coordinate)
!      call set_param_value(at, trim(calc_energy_per_coordinate), energy_per_coordinate)
!   end if
!   if (do_local_gap_variance) then                           ! :558-571 (sum_in_place over mpi first, :545-548)
!      call add_property(at, trim(calc_local_gap_variance), 0.0_dp, ptr=local_gap_variance_pointer)
!      call add_property(at, "gap_variance_gradient", 0.0_dp, n_cols=3, ptr2=gap_variance_gradient_pointer)
!      ierr = gap_potential_get_local_gap_variance(this%b200, at%N, local_gap_variance_pointer, c_loc(gap_variance_gradient_pointer))
!   end if
!   ! IPModel_GAP.f95:538-556 (sum_in_place(mpi, f / virial / local_virial / local_e), e = sum(mpi, e)) is GONE: with the communicator
!   ! set, e, f, virial, local_e and local_virial come back from gap_potential_calc already summed over the ranks (one reduction of
!   ! the packed [E | virial | F] buffer on the GPUs, over NVLink).  Only the optional energy_per_coordinate / local_gap_variance
!   ! arrays are still per-rank partial sums and keep their sum_in_place calls (:545-549).
! end subroutine
