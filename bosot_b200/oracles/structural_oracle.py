"""A deterministic black-box objective over kernel expressions, for examples and tests.

The reference's objectives (Laplace model evidence / BIC / CV of a GP fitted to data,
bosot/oracles/*.py) are a different workload (dense GP fits on n <= 500 data points through GPflow)
and out of scope here; BO only needs ``query(expr) -> (value, seconds)`` (base_object_oracle.py).
This oracle scores an expression by how well its structure matches a hidden target kernel: a sum
of per-symbol affinities, a depth penalty, and a bonus for containing a target sub-structure --
cheap, deterministic, and with enough structure for the surrogate to learn from."""
import time

import numpy as np


class StructuralOracle:
    def __init__(self, symbols, seed=0, noise=0.0):
        rng = np.random.default_rng(seed)
        self.affinity = {s: float(a) for s, a in zip(symbols, rng.normal(0.0, 1.0, size=len(symbols)))}
        self.pair_bonus = {}
        for _ in range(3):
            a, b = rng.choice(len(symbols), size=2, replace=False)
            op = "ADD" if rng.integers(2) == 0 else "MULTIPLY"
            self.pair_bonus[(op, frozenset((symbols[a], symbols[b])))] = float(rng.uniform(0.5, 1.5))
        self.noise = noise

    def _score(self, e):
        op = e.get_operator()
        if op is None:
            return self.affinity[e.get_name()], 1, e.get_name()
        s1, n1, k1 = self._score(e.expression1)
        s2, n2, k2 = self._score(e.expression2)
        s = s1 + s2 if op.name == "ADD" else 0.5 * (s1 + s2) + 0.25 * s1 * s2
        if k1 is not None and k2 is not None:
            s += self.pair_bonus.get((op.name, frozenset((k1, k2))), 0.0)
        return s, n1 + n2, None

    def query(self, x):
        t0 = time.perf_counter()
        s, n_leaves, _ = self._score(x)
        value = s / np.sqrt(n_leaves) - 0.05 * n_leaves
        if self.noise:
            value += float(np.random.normal(0, self.noise))
        return value, time.perf_counter() - t0
