"""Value-network ensembles as costate laws for the forward closed-loop simulations (levelsets.py:726-933,
eval_utils.py:10-45; nn_utils.py:119-199).

The reference trains an ensemble of flax MLPs (`my_nn_flax`: Dense -> softplus per hidden layer, Dense(1)) on
NORMALISED data (`data_normaliser`) and drives forward simulations with lambda(x) = d/dx mean_e V_e(x),
V_e(x) = unnormalise_v(nn(params_e, normalise_x(x))).  Training is out of scope here; this module takes the
trained parameters in flax's own layout and hands them to the CUDA kernels:

    params = {'params': {'Dense_0': {'kernel': (E, nx, W), 'bias': (E, W)}, 'Dense_1': {'kernel': (E, W, W), ...}, ...,
                         'Dense_L': {'kernel': (E, W, 1), 'bias': (E, 1)}}}          # leading ensemble axis (vmapped init)
    ens = NnEnsemble.from_flax(params, normaliser, device, dtype)
    lam, v_mean, v_std = ens.meanstd(x)               # vx of the mean / v_meanstds, levelsets.py:789-815
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _capi

__all__ = ["DataNormaliser", "NnEnsemble", "init_flax_params"]


class DataNormaliser:
    """nn_utils.data_normaliser (:119-152): zero mean / unit variance of every x component and of v."""

    def __init__(self, train_ode_states=None, *, x_means=None, x_stds=None, v_mean=0.0, v_std=1.0):
        if train_ode_states is not None:
            x = np.asarray(_host(train_ode_states["x"]), dtype=np.float64)
            v = np.asarray(_host(train_ode_states["v"]), dtype=np.float64)
            x_means, x_stds, v_mean, v_std = x.mean(0), x.std(0), v.mean(), v.std()
        self.x_means = np.asarray(x_means, dtype=np.float64)
        self.x_stds = np.asarray(x_stds, dtype=np.float64)
        self.v_mean, self.v_std = float(v_mean), float(v_std)

    @classmethod
    def identity(cls, nx):
        return cls(x_means=np.zeros(nx), x_stds=np.ones(nx))

    def normalise_x(self, x): return (x - self.x_means) / self.x_stds
    def unnormalise_x(self, xn): return xn * self.x_stds + self.x_means
    def normalise_v(self, v): return (v - self.v_mean) / self.v_std
    def unnormalise_v(self, vn): return vn * self.v_std + self.v_mean
    def normalise_vx(self, vx): return vx * self.x_stds / self.v_std
    def unnormalise_vx(self, vxn): return (vxn / self.x_stds) * self.v_std


def _host(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def init_flax_params(seed: int, nx: int, layer_dims=(64, 64, 64), ensemble: int = 8, scale: float = 1.0) -> dict:
    """Random parameters with flax's default initialisers (lecun_normal kernels, zero biases) in flax's layout, with a
    leading ensemble axis.  For tests and benchmarks: the reference's weights come out of its training loop."""
    rng = np.random.Generator(np.random.PCG64(seed))
    dims = [nx, *layer_dims, 1]
    out = {}
    for i in range(len(dims) - 1):
        fan_in = dims[i]
        k = rng.standard_normal((ensemble, dims[i], dims[i + 1])) * scale / np.sqrt(fan_in)
        out[f"Dense_{i}"] = {"kernel": k, "bias": 0.1 * rng.standard_normal((ensemble, dims[i + 1]))}
    return {"params": out}


@dataclass
class NnEnsemble:
    weights: torch.Tensor      # flat device array in the layout of pmp_nn_costate_t
    nx: int
    width: int
    n_hidden: int
    n_nets: int
    normaliser: DataNormaliser
    n_sigma: float = 2.0       # levelsets.py:908
    sigma_max: float = 0.5     # levelsets.py:909

    @classmethod
    def from_flax(cls, params: dict, normaliser: DataNormaliser, device, dtype=torch.float32, **kw) -> "NnEnsemble":
        layers = params["params"] if "params" in params else params
        names = sorted(layers, key=lambda s: int(s.split("_")[-1]))
        ks = [np.asarray(_host(layers[n]["kernel"]), dtype=np.float64) for n in names]
        bs = [np.asarray(_host(layers[n]["bias"]), dtype=np.float64) for n in names]
        if ks[0].ndim == 2:  # a single network: ensemble of one
            ks, bs = [k[None] for k in ks], [b[None] for b in bs]
        E, nx, W = ks[0].shape
        L = len(ks) - 1
        if not all(k.shape == (E, W, W) for k in ks[1:-1]) or ks[-1].shape != (E, W, 1):
            raise NotImplementedError("NnEnsemble: hidden layers must share one width and the output must be scalar")
        if W not in (32, 64) or not 1 <= L <= 4 or not 1 <= E <= 16:
            raise NotImplementedError("NnEnsemble: width in {32, 64}, 1..4 hidden layers, 1..16 members")
        rows = []
        for e in range(E):
            parts = [ks[0][e].reshape(-1), bs[0][e]]
            for l in range(1, L):
                parts += [ks[l][e].reshape(-1), bs[l][e]]
            parts += [ks[L][e].reshape(-1), bs[L][e].reshape(-1), np.zeros(3)]
            rows.append(np.concatenate(parts))
        flat = np.stack(rows)
        assert flat.shape[1] == nx * W + W + (L - 1) * (W * W + W) + W + 4
        w = torch.as_tensor(flat.reshape(-1), dtype=dtype, device=device).contiguous()
        return cls(w, nx, W, L, E, normaliser, **kw)

    def to_c(self) -> _capi.NnCostate:
        c = _capi.NnCostate(n_nets=self.n_nets, n_hidden_layers=self.n_hidden, width=self.width, reserved=0,
                            weights=self.weights.data_ptr(), v_mean=self.normaliser.v_mean, v_std=self.normaliser.v_std,
                            n_sigma=self.n_sigma, sigma_max=self.sigma_max)
        for q in range(16):
            c.x_mean[q] = float(self.normaliser.x_means[q]) if q < self.nx else 0.0
            c.x_std[q] = float(self.normaliser.x_stds[q]) if q < self.nx else 1.0
        return c

    def meanstd(self, problem: _capi.Problem, x):
        """(vx of the ensemble mean, v_mean, v_std) at x (nx,) or (N, nx): v_meanstds / vx_meanstd, levelsets.py:789-815."""
        dev, dtype = self.weights.device, self.weights.dtype
        xt = torch.as_tensor(_host(x) if not isinstance(x, torch.Tensor) else x, dtype=dtype, device=dev)
        batched = xt.dim() == 2
        xs = xt.reshape(-1, self.nx).t().contiguous()
        n = xs.shape[1]
        lam = torch.empty_like(xs)
        vm, vs = torch.empty(n, dtype=dtype, device=dev), torch.empty(n, dtype=dtype, device=dev)
        L = _capi.lib_of(problem)
        c = self.to_c()
        with torch.cuda.device(dev):
            _capi.check(L.pmp_nn_value_batch_cuda(C.byref(problem), C.byref(c), _capi.F64 if dtype == torch.float64 else _capi.F32,
                                                  n, xs.data_ptr(), lam.data_ptr(), vm.data_ptr(), vs.data_ptr(),
                                                  C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), L)
        lam = lam.t()
        return (lam, vm, vs) if batched else (lam[0], vm[0], vs[0])
