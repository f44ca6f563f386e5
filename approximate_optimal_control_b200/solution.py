"""`Solution`: the subset of diffrax.Solution that the reference's callers touch, backed by torch
CUDA tensors.

Attributes used by the reference (file:line in /root/reference):
  .ts      (N, S+1)            saved times, padded -inf/+inf      visualiser.py:97-99
  .ys      {'x','t','v','vx'}  (N, S+1[, nx]), padded +inf        levelsets.py:679-693, visualiser.py:76-77
  .stats   {'num_steps','num_accepted_steps','num_rejected_steps'} levelsets.py:1215, rrt_sampler.py:497
  .result  (N,) int32, 1 == terminating event                     levelsets.py:1218
  .t0 .t1
Extension: .y_end = {'x','t','v','vx'} the last accepted state, i.e. ys[..][num_accepted_steps]
(levelsets.py:1215) without materialising the dense buffers.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import torch


class StepStats(dict):
    """diffrax's `stats` dict; 'num_rejected_steps' = num_steps - num_accepted_steps is formed on first access (on host
    tensors of 2^20 entries the subtraction is a measurable part of a 2 ms solve, and the reference's callers never read it)."""

    def __init__(self, num_steps, num_accepted_steps):
        super().__init__(num_steps=num_steps, num_accepted_steps=num_accepted_steps)

    def __missing__(self, key):
        if key != "num_rejected_steps":
            raise KeyError(key)
        self[key] = self["num_steps"] - self["num_accepted_steps"]
        return self[key]

    def _full(self):
        self["num_rejected_steps"]
        return self

    def keys(self):
        return self._full() and super().keys()

    def items(self):
        return self._full() and super().items()

    def values(self):
        return self._full() and super().values()

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return 3

    def __contains__(self, key):
        return key in ("num_steps", "num_accepted_steps", "num_rejected_steps")


@dataclass
class Solution:
    t0: float
    t1: float
    ts: Optional[torch.Tensor]
    ys: Optional[Dict[str, torch.Tensor]]
    y_end: Dict[str, torch.Tensor]
    stats: Dict[str, torch.Tensor]
    result: torch.Tensor
    batched: bool = True
    extra: dict = field(default_factory=dict)

    def evaluate(self, t):
        """diffrax Solution.evaluate (SaveAt(dense=True)): the solver's interpolant at time(s) `t`
        (value level(s) for a value-parameterised solution).  Batched solution: `t` is a scalar, (N,) or
        (N, M); single solution: scalar or (M,) -- the latter replaces jax.vmap(sol.evaluate)(ts)
        (experiment_orbits.py:183-184).  Queries outside the integrated interval give NaN."""
        fn = self.extra.get("interp")
        if fn is None or self.ys is None:
            raise RuntimeError("evaluate() needs the saved steps: solve with save_steps=True")
        return fn(self, t, 0)

    def state_at_value(self, v_k):
        """State on the trajectory at which v == v_k (the bisection on sol.evaluate of misc.py:47-91), found
        inside the accepted step that brackets the level.  Time-parameterised solutions only."""
        fn = self.extra.get("interp")
        if fn is None or self.ys is None:
            raise RuntimeError("state_at_value() needs the saved steps: solve with save_steps=True")
        return fn(self, v_k, 1)

    def numpy(self) -> "Solution":
        """Copy every tensor to host numpy arrays (same structure)."""
        cv = lambda t: None if t is None else t.detach().cpu().numpy()
        cd = lambda d: ({k: cv(v) for k, v in d.items()} if isinstance(d, dict) else cv(d))
        return Solution(self.t0, self.t1, cv(self.ts), cd(self.ys), cd(self.y_end), cd(self.stats),
                        cv(self.result), self.batched, dict(self.extra))

    def __getitem__(self, i) -> "Solution":
        """Select trajectories (the reference does jax.tree_util.tree_map(itemgetter(i), sols))."""
        sel = lambda t: None if t is None else t[i]
        sd = lambda d: ({k: v[i] for k, v in d.items()} if isinstance(d, dict) else sel(d))
        return Solution(self.t0, self.t1, sel(self.ts), sd(self.ys), sd(self.y_end), sd(self.stats),
                        sel(self.result), not isinstance(i, int), dict(self.extra))
