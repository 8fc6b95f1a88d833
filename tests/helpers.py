"""Shared test helpers.  The oracle is imported here as the CHECKER only."""
import numpy as np

from oracle import OracleParams, oracle_consensus_batch, pack_groups  # noqa: F401


class OracleBackedContext:
    """Test double with PoaContext's upload/run/fetch surface, backed by the oracle.
    Lets the host logic (batching, ordering, fall-backs, sharding) be tested without a GPU."""

    def __init__(self, params=None):
        self.params = params
        self._packed = None
        self._flags = None
        self._out = None
        self.calls = 0

    def upload(self, gro, rbo, bases, flags=None):
        self._packed = (np.asarray(gro, np.int64), np.asarray(rbo, np.int64), np.asarray(bases, np.uint8))
        self._flags = None if flags is None else np.asarray(flags, np.uint8)

    def run(self):
        self.calls += 1
        self._out = oracle_consensus_batch(packed=self._packed, params=self.params, flags=self._flags)
        st = dict(self._out["stats"])
        st.setdefault("kernel_ms", 1.0)
        st["kernel_ms"] = st["kernel_ms"] or 1.0
        return st

    def fetch(self, trace=False):
        return dict(cons=self._out["cons"], status=self._out["status"], trace=None)

    def consensus_batch(self, groups=None, packed=None, trace=False, flags=None):
        self.upload(*(packed if packed is not None else pack_groups(groups)), flags=flags)
        st = self.run()
        out = self.fetch()
        out["stats"] = st
        return out


def random_seq(rng, n):
    return "".join(rng.choice(list("ACGT"), size=n))
