"""climatemachine.jl_b200 -- B200-native DG tendency + LSRK path behind ClimateMachine.jl's
DGModel / MPIStateArray / BalanceLaw interface.

The compute path is the hand-written sm_100a CUDA library ``libcmdg.so`` (csrc/, C ABI in
include/cmdg.h).  This Python package is the host-side harness that mirrors the reference's
interface for the path (same names, argument meaning and error behaviour); in production the
host is Julia and binds the same C ABI with ``ccall``.
"""
from . import _lib
from .balance_laws import *  # noqa: F401,F403
from .balance_laws import UnsupportedModelError
from .dgmodel import (DGModel, DiscontinuousSpectralElementGrid, MPIStateArray,
                      LowStorageRungeKutta2N, LSRK54CarpenterKennedy,
                      LSRK144NiegemannDiehlBusch, solve, norm, euclidean_distance,
                      comm_unique_id, ErrorOnRemoteNode)

__all__ = [n for n in dir() if not n.startswith("_")]
