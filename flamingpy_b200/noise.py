"""CV noise layer: reference-compatible ``CVLayer`` plus the host-side samplers.

``CVLayer`` keeps the reference's constructor and ``apply_noise(rng)`` contract
(``flamingpy/noise/cv.py:27-119``): with a NumPy generator it draws the labels
and quadratures on the host in the reference's own RNG call order (so a shared
seed gives the same realisation as the reference) and pushes them through the
CUDA pipeline as an injected trial; the batched Monte-Carlo driver
(``simulations.ec_monte_carlo``) instead uses the device Philox stream.
"""

from __future__ import annotations

from typing import Optional

import numpy as np


class LabelSampler:
    """State labels of a (reused) layer, ``CVLayer.populate_states``
    (``noise/cv.py:121-140,306-321``).

    ``p_swap`` truthy: ``num_p = rng.binomial(N, p_swap)`` then
    ``rng.choice(range(N), num_p, replace=False)`` (``p_swap == 1``: all nodes, no
    RNG call).  The GKP set is every node in neither ``states["p"]`` nor the
    *previous* call's ``states["GKP"]`` - the reference never clears that entry -
    so a reused layer leaves some nodes noiseless (SURVEY.md finding 0.3);
    ``fresh=True`` models a new layer per trial instead.
    Codes: 0 silent, 1 noisy GKP, 2 p-squeezed.
    """

    def __init__(self, n_nodes: int, p_swap, fresh: bool = False):
        self.N = int(n_nodes)
        self.p_swap = p_swap
        self.fresh = fresh
        self._prev_gkp = np.zeros(self.N, bool)

    def next(self, rng) -> np.ndarray:
        N = self.N
        is_p = np.zeros(N, bool)
        if self.p_swap:
            if self.p_swap == 1:
                is_p[:] = True
            else:
                num_p = rng.binomial(N, self.p_swap)
                is_p[rng.choice(range(N), size=int(np.floor(num_p)), replace=False)] = True
        gkp = ~is_p if self.fresh else (~is_p & ~self._prev_gkp)
        self._prev_gkp = gkp
        state = np.zeros(N, np.uint8)
        state[gkp] = 1
        state[is_p] = 2
        return state


def sigma_vector(state: np.ndarray, delta: float) -> np.ndarray:
    """float32[2N] standard deviations of the "initial" sampling order,
    ``CVLayer._covs_sampler`` (``noise/cv.py:277-290``)."""
    N = state.shape[-1]
    sig = np.zeros(state.shape[:-1] + (2 * N,), dtype=np.float32)
    gkp_sd = (delta / 2) ** 0.5
    p_q_sd = 1 / (2 * delta) ** 0.5
    q, p = sig[..., :N], sig[..., N:]
    q[state == 1] = gkp_sd
    q[state == 2] = p_q_sd
    p[state != 0] = gkp_sd
    return sig


def draw(rng, sigma: np.ndarray) -> np.ndarray:
    """``rng.normal(means, covs)`` with float32 zero means (``noise/cv.py:205``)."""
    return rng.normal(np.zeros(sigma.shape, dtype=np.float32), sigma)


def sample_batch(rng, n_nodes: int, delta: float, p_swap, n_trials: int, fresh: bool = False,
                 sampler: Optional[LabelSampler] = None, out_x=None, out_state=None):
    """Host-side labels + draws for ``n_trials`` consecutive trials of one chain."""
    sampler = sampler or LabelSampler(n_nodes, p_swap, fresh)
    x = out_x if out_x is not None else np.empty((n_trials, 2 * n_nodes), np.float64)
    st = out_state if out_state is not None else np.empty((n_trials, n_nodes), np.uint8)
    for t in range(n_trials):
        st[t] = sampler.next(rng)
        x[t] = draw(rng, sigma_vector(st[t], delta))
    return x, st
