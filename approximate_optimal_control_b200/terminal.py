"""Terminal-set sampling: the caller block of the hot path (levelsets.py:540-599).

Samples points on the LQR value level set {x : 1/2 (x-x_eq)' P (x-x_eq) = V_f} and attaches the
terminal value and costate of the LQR value function, v = V_f, lambda = P (x - x_eq), t = 0.
"""
from __future__ import annotations

import numpy as np


def sample_terminal_states(P_lqr, V_f: float, n: int, seed: int = 1, x_eq=None, dtype=np.float64):
    """levelsets.py:547-556: normal points -> unit sphere -> x_f = s @ inv(L) * sqrt(V_f) * sqrt(2)
    with P = L L'.  jax's PRNGKey stream is not reproducible without jax; numpy's PCG64(seed) is used
    (the inputs of the CUDA path and of the oracle are the same arrays, which is what parity needs).

    Returns the reference's state dict with a leading batch axis:
        {'x': (n,nx), 't': (n,), 'v': (n,), 'vx': (n,nx)}
    """
    P = np.asarray(P_lqr, dtype=np.float64)
    nx = P.shape[0]
    L = np.linalg.cholesky(P)
    assert np.linalg.norm(L @ L.T - P) / np.linalg.norm(P) < 1e-6, "cholesky decomposition wrong or inaccurate"
    rng = np.random.Generator(np.random.PCG64(seed))
    normal_pts = rng.standard_normal((n, nx))
    unitsphere_pts = normal_pts / np.linalg.norm(normal_pts, axis=1)[:, None]
    dx = unitsphere_pts @ np.linalg.inv(L) * np.sqrt(V_f) * np.sqrt(2)
    v = 0.5 * np.einsum("ni,ij,nj->n", dx, P, dx)
    assert np.allclose(v, V_f), "wrong terminal value..."  # levelsets.py:562
    x = dx if x_eq is None else dx + np.asarray(x_eq, dtype=np.float64)[None]
    return {"x": x.astype(dtype), "t": np.zeros(n, dtype), "v": v.astype(dtype), "vx": (dx @ P.T).astype(dtype)}


def ellipse_terminal_states(P0, n: int, radius: float, x_eq=None, dtype=np.float64):
    """Terminal states of the two 2-state experiments (experiment_orbits.py:121-129,
    experiment_lqr_statepenalty.py:123-129): n points on the ellipse
        x_f = x_eq + radius * sqrtm(P0)^-1 (sin th, cos th),  th = 2 pi i / n,
    with the terminal value function V(x) = (x-x_eq)' P0 (x-x_eq): v_f = V(x_f), lambda_f = 2 P0 (x_f - x_eq).
    Returns the state dict with a leading batch axis."""
    import scipy.linalg
    P0 = np.asarray(P0, dtype=np.float64)
    th = 2 * np.pi * np.arange(n) / n
    circ = np.stack([np.sin(th), np.cos(th)], axis=1)                       # (n, 2)
    dx = radius * circ @ np.linalg.inv(scipy.linalg.sqrtm(P0).real).T
    v = np.einsum("ni,ij,nj->n", dx, P0, dx)
    x = dx if x_eq is None else dx + np.asarray(x_eq, dtype=np.float64)[None]
    return {"x": x.astype(dtype), "t": np.zeros(n, dtype), "v": v.astype(dtype), "vx": (2 * dx @ P0.T).astype(dtype)}
