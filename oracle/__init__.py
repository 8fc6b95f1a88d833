"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/mpoa_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  PARITY UNPINNED: see the header of oracle/abpoa_oracle.cpp.
"""
from .pyoracle import OracleParams, oracle_consensus_batch, build_oracle, pack_groups  # noqa: F401
