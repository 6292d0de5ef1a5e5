"""CPU oracle for the ClimateMachine.jl DG tendency + LSRK hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a NumPy (and, under ``oracle/c``, plain C: ``dg_ref.c`` for the dry AtmosModel, ``hb_ref.c`` for the
ocean HBModel, both checked against the NumPy code in ``tests/test_oracle_c.py``) restatement of the reference
algorithm (CliMA/ClimateMachine.jl v0.3.0-DEV, pure Julia) for the single hot
path named by BASELINE.json.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it,
and there only as the *checker* (or the CPU arm being timed) -- never on the
product path.  The product path (``climatemachine.jl_b200``) fails loudly when
its CUDA library is missing; it never falls back to this package.

The reference cannot be executed in this environment (no Julia, no network), so
the oracle is pinned against the reference's *own* golden numbers instead:

* ``test/Numerics/DGMethods/Euler/isentropicvortex.jl:60-238`` -- L2 error table
  (mesh + metrics + DGModel + Rusanov/Central/Roe + LSRK54 + thermodynamic
  constants, end to end), reproduced to ``rtol = sqrt(eps)`` as the test itself
  demands (``tests/test_oracle_golden.py``: levels 1-2 in NumPy; levels 3-4, Rusanov and Central, by the C twin
  in ``tests/test_oracle_c.py``);
* ``test/Numerics/DGMethods/compressible_Navier_Stokes/mms_bc_atmos.jl`` -- the second-order
  (Navier-Stokes) path: 3-D level-1 golden error to 1e-7 (``tests/test_oracle_mms.py``);
* ``test/Numerics/DGMethods/advection_diffusion/periodic_3D_hyperdiffusion.jl`` -- the
  hyperdiffusion kernels: 3-D HorizontalDirection level-1 golden error to 1e-7
  (``tests/test_oracle_hyperdiffusion_golden.py``);
* ``test/Numerics/DGMethods/grad_test_sphere.jl``, ``grad_test.jl`` -- cubed-sphere metric terms and the
  element-local gradient (``tests/test_oracle_grad.py``);
* ``test/Atmos/Model/discrete_hydrostatic_balance.jl`` -- Gravity + HydrostaticState stay balanced to
  100 eps on the LES box and the GCM shell (``tests/test_oracle_balance.py``);
* ``test/Numerics/DGMethods/courant.jl``, ``test/Numerics/Mesh/min_node_distance.jl`` -- Courant numbers
  and node distances (``tests/test_oracle_courant.py``);
* ``test/Numerics/Mesh/filter.jl`` -- golden filter matrices and the analytic application test
  (``tests/test_oracle_filters.py``, fixtures in ``tests/golden``);
* ``test/Ocean/refvals/test_ocean_gyre_refvals.jl`` (short) and ``test_windstress_refvals.jl``
  (explicit_cpu) -- HBModel regression values (``tests/test_oracle_ocean.py``,
  ``tests/test_oracle_ocean_windstress.py``); ``test/Ocean/HydrostaticBoussinesq/test_3D_spindown.jl`` with
  ``refvals/3D_hydrostatic_spindown_refvals.jl`` (explicit) -- 720 LSRK144 steps of the SimpleBox spin-down
  (periodic, free-slip / penetrable free-slip boundaries) by the C twin: statistics to 1.5e-12, and the error
  against the analytic solution equal to the value the reference prints to 1e-14
  (``tests/test_oracle_ocean_spindown.py``); ``test/Numerics/DGMethods/integral_test.jl`` -- the stack
  integral (``tests/test_oracle_integral.py``);
* ``test/Numerics/ODESolvers/ode_tests_convergence.jl`` -- LSRK54 / LSRK144 order 4 on the reference's
  time-dependent problem;
* ``test/Numerics/Mesh/mpi_connect*.jl`` -- connectivity / ghost lists on 3-5 ranks
  (bit exact);
* ``test/Arrays/mpi_comm.jl`` -- halo pack/unpack known answers;
* ``test/Numerics/Mesh/{Elements,Grids,Metrics}.jl`` -- quadrature / metric identities.

Parts for which the reference holds no tight golden value are "parity unpinned" (restatement plus
analytic / property checks only), here and in DESIGN.md section 3: Coriolis, the Smagorinsky-Lilly
closure itself (its plumbing is pinned by the MMS run), HeldSuarezForcing and RayleighSponge, the
AtmosModel hooks of DryBiharmonic (its kernels are pinned), NTracers, the rising-bubble configuration
with DryAdiabaticProfile, the AtmosFilterPerturbations target.

Array convention: every array is stored with the *bytes* Julia would have
(column-major ``A[i, j, k]`` == C-order ``a[k, j, i]``), so a NumPy array of
shape ``(nelem, nstate, Np)`` is byte-identical to the reference's
``Np x nstate x nelem`` ``MPIStateArray.data``.  Index arrays (``vmapM``,
``vmapP``, ``vmapsend`` ...) hold the reference's 1-based Int64 values.
"""
