"""CPU oracle for the ClimateMachine.jl DG tendency + LSRK hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a NumPy (and, under ``oracle/c``, plain C) restatement of the reference
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
  demands (``tests/test_oracle_golden.py``);
* ``test/Numerics/Mesh/mpi_connect*.jl`` -- connectivity / ghost lists on 3-5 ranks
  (bit exact);
* ``test/Arrays/mpi_comm.jl`` -- halo pack/unpack known answers;
* ``test/Numerics/Mesh/{Elements,Grids,Metrics}.jl`` -- quadrature / metric identities.

Parts for which the reference holds no tight golden value (Coriolis/gravity on
the cubed sphere, Smagorinsky on the sphere) are marked "parity unpinned" in
the module that restates them and in DESIGN.md.

Array convention: every array is stored with the *bytes* Julia would have
(column-major ``A[i, j, k]`` == C-order ``a[k, j, i]``), so a NumPy array of
shape ``(nelem, nstate, Np)`` is byte-identical to the reference's
``Np x nstate x nelem`` ``MPIStateArray.data``.  Index arrays (``vmapM``,
``vmapP``, ``vmapsend`` ...) hold the reference's 1-based Int64 values.
"""
