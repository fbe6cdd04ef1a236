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


class CVLayer:
    """Reference-compatible CV noise layer (``flamingpy/noise/cv.py:27-349``).

    ``CVLayer(code, *, delta, p_swap=None, states=None, sampling_order="initial",
    translator=None)``; ``apply_noise(rng)`` labels the nodes, samples the
    quadratures on the host with the reference's RNG call order, and runs the
    entangle / bin stage on the GPU, writing ``state``, ``hom_val_p`` and
    ``bit_val`` back to ``code.graph``.  A reused layer carries the previous
    trial's GKP set over exactly like the reference (SURVEY.md finding 0.3).
    """

    def __init__(self, code, *, delta, **kwargs):
        from .codes import SurfaceCode

        if not isinstance(code, SurfaceCode):
            raise TypeError("CVLayer needs a flamingpy_b200.codes.SurfaceCode")
        self.code = code
        self.egraph = code.graph
        self._N = len(code.graph)
        self._to_points = code.graph.to_points
        supplied_states = kwargs.get("states")
        self.delta = delta
        self.p_swap = kwargs.get("p_swap")
        self.states = supplied_states or {"p": np.empty(0, dtype=int)}
        self._sampling_order = kwargs.get("sampling_order") or "initial"
        if self._sampling_order != "initial":
            raise NotImplementedError("only the 'initial' sampling order is live for CVLayer (SURVEY.md 8c)")
        if kwargs.get("translator") is not None:
            raise NotImplementedError("the GPU path implements the standard GKP binner only")
        if self.p_swap is not None and supplied_states is not None and len(supplied_states["p"]):
            raise Exception(
                "Both the swap-out probability and indices of p-squeezed states "
                "have been supplied. Please only specify one of those."
            )
        self._sampler = LabelSampler(self._N, self.p_swap)
        self._init_covs = None

    # -- reference API ---------------------------------------------------------------
    def apply_noise(self, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        self.populate_states(rng)
        self.measure_syndrome(rng)
        self.inner_decoder()

    def populate_states(self, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        if self.p_swap:
            state = self._sampler.next(rng)
        else:
            # indices supplied in states["p"] (or none): same carry-over of states["GKP"]
            is_p = np.zeros(self._N, bool)
            is_p[np.asarray(self.states.get("p", []), int)] = True
            gkp = ~is_p & ~self._sampler._prev_gkp
            self._sampler._prev_gkp = gkp
            state = np.zeros(self._N, np.uint8)
            state[gkp] = 1
            state[is_p] = 2
        self.states["p"] = np.nonzero(state == 2)[0]
        self.states["GKP"] = np.nonzero(state == 1)[0]
        g = self.egraph
        g._ensure_arrays()
        # nodes in neither set keep their stale label: anything not "p" reads "GKP"
        g.state = state.copy()
        self._state = state

    def measure_syndrome(self, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        sig = sigma_vector(self._state, self.delta)
        self._init_covs = sig
        x = draw(rng, sig)
        self.code._trial = (x, self._state.copy(), self.delta, self.p_swap)
        eng = self.code.engine(self.delta, self.p_swap)
        from . import _lib

        batch = eng.run_injected(x[None, :], self._state[None, :], stage=_lib.STAGE_NOISE)
        g = self.egraph
        off = 0
        for e in self.code.ec:
            ct = self.code.lattice.complexes[e]
            g.hom_val_p[ct.qubits] = batch.hom[0, off : off + ct.Q]
            g.bit_val[ct.qubits] = batch.bits[0, off : off + ct.Q]
            off += ct.Q

    def inner_decoder(self):
        """Binning already happened on the device together with the homodyne step."""

    def hom_outcomes(self, inds=None, quad="p"):
        if quad != "p":
            raise NotImplementedError("only p-homodyne outcomes are produced on this path")
        inds = range(self._N) if inds is None else inds
        h = self.egraph.hom_val_p
        return [None if (h is None or np.isnan(h[i])) else h[i] for i in inds]

    def bit_values(self, inds=None):
        inds = range(self._N) if inds is None else inds
        b = self.egraph.bit_val
        return [None if (b is None or b[i] < 0) else int(b[i]) for i in inds]

    @property
    def p_inds(self):
        return self.states.get("p")

    @property
    def gkp_inds(self):
        return self.states.get("GKP")


class MacroSampler:
    """The random draws of ``CVMacroLayer.apply_noise`` on a (reused) layer, in the reference's RNG
    call order (``noise/cv.py:375-403``), for the 4N micronodes of the macronode lattice:

    1. labels as in ``LabelSampler`` (``cv.py:121-140,306-321``), over the micronodes;
    2. the initial q means of the two-step sampling (``cv.py:323-349``): ``rng.random(n_p) * 2 sqrt(pi)``
       for the p-squeezed micronodes in the order ``rng.choice`` returned them, then
       ``rng.integers(0, 2, n_gkp) * sqrt(pi)`` for the GKP micronodes in the order of the reference's
       ``list(set(range(N)) - set(used))`` - a Python-set order, reproduced with a real set - both
       stored as float32; micronodes in neither set (reused layer) keep mean 0;
    3. standard normals for the N star p-measurements, then the 3N planet q-measurements
       (``rng.normal(means, covs)`` is ``means + covs * standard_normal`` element by element).
    """

    def __init__(self, n_micro: int, p_swap, fresh: bool = False):
        self.M = int(n_micro)
        self.p_swap = p_swap
        self.fresh = fresh
        self._prev_gkp = np.empty(0, dtype=int)

    def next(self, rng):
        """Returns (state uint8[M], mean32 float32[M], z float64[M])."""
        M = self.M
        if self.p_swap:
            if self.p_swap == 1:
                p_inds = np.arange(M)
            else:
                num_p = rng.binomial(M, self.p_swap)
                p_inds = rng.choice(range(M), size=int(np.floor(num_p)), replace=False)
        else:
            p_inds = np.empty(0, dtype=int)
        used = np.concatenate([p_inds, np.empty(0, dtype=int) if self.fresh else self._prev_gkp])
        gkp_inds = np.array(list(set(range(M)) - set(used)), dtype=int)
        self._prev_gkp = gkp_inds
        state = np.zeros(M, np.uint8)
        state[gkp_inds] = 1
        state[p_inds] = 2
        mean = np.zeros(M, np.float32)
        if len(p_inds):
            mean[p_inds] = rng.random(size=len(p_inds)) * (2 * np.sqrt(np.pi))
        if len(gkp_inds):
            mean[gkp_inds] = rng.integers(0, 2, size=len(gkp_inds)) * np.sqrt(np.pi)
        z = np.concatenate([rng.standard_normal(M // 4), rng.standard_normal(M - M // 4)])
        return state, mean, z


class CVMacroLayer:
    """Reference-compatible macronode noise layer (``flamingpy/noise/cv.py:352-625``; the reference
    CLI's default noise model, ``simulations.py:224``).

    ``CVMacroLayer(code, *, delta, p_swap=None)``; ``apply_noise(rng)`` draws the micronode labels,
    initial means and measurement noise on the host in the reference's RNG call order, and the GPU
    entangles the dumbbells, applies the four-splitter of every macronode, measures stars in p and
    planets in q and reduces each macronode to the canonical lattice: ``state``, ``bit_val`` and
    ``p_phase_cond`` of every node of ``code.graph``.  Decode with
    ``weight_options={"method": "blueprint", "prob_precomputed": True}`` (``decoder.py:67-68``).
    """

    def __init__(self, code, *, delta, bs_network=None, **kwargs):
        from .codes import SurfaceCode

        if not isinstance(code, SurfaceCode):
            raise TypeError("CVMacroLayer needs a flamingpy_b200.codes.SurfaceCode")
        if bs_network is not None:
            raise NotImplementedError("only the standard four-splitter (cv/ops.py:104-127) is implemented")
        if kwargs.get("states") is not None and len(kwargs["states"].get("p", [])):
            raise NotImplementedError("CVMacroLayer with hand-picked p-squeezed micronodes: use p_swap")
        if kwargs.get("translator") is not None:
            raise NotImplementedError("the GPU path implements the standard GKP binner only")
        self.code = code
        self.reduced_graph = code.graph
        self.delta = delta
        self.p_swap = kwargs.get("p_swap")
        self._N = 4 * len(code.graph)  # micronodes
        self._sampler = MacroSampler(self._N, self.p_swap)
        self.states = {"p": np.empty(0, dtype=int)}

    def apply_noise(self, rng=None):
        from . import _lib

        rng = np.random.default_rng() if rng is None else rng
        state, mean, z = self._sampler.next(rng)
        self.states = {"p": np.nonzero(state == 2)[0], "GKP": np.nonzero(state == 1)[0]}
        code = self.code
        code._trial = ("macro", (state, mean, z), self.delta, self.p_swap)
        eng = code.engine(self.delta, self.p_swap, {"method": "blueprint", "prob_precomputed": True})
        red = eng.run_macro_injected(state[None, :], mean[None, :], z[None, :], stage=_lib.STAGE_NOISE)
        g = self.reduced_graph
        g._ensure_arrays()
        g.state = red["red_state"][0].copy()
        g.bit_val[:] = red["bit_val"][0]
        g.p_phase_cond = red["p_phase_cond"][0].copy()
        g.hom_val_p[:] = np.nan  # the reduced lattice carries no homodyne values (cv.py:622-625)

    def bit_values(self, inds=None):
        inds = range(len(self.code.graph)) if inds is None else inds
        b = self.reduced_graph.bit_val
        return [None if (b is None or b[i] < 0) else int(b[i]) for i in inds]

    @property
    def p_inds(self):
        return self.states.get("p")

    @property
    def gkp_inds(self):
        return self.states.get("GKP")


class IidNoise:
    """Reference-compatible i.i.d. Z-error model (``flamingpy/noise/iid.py:19-50``):
    ``IidNoise(code, error_probability).apply_noise(rng)`` sets ``bit_val = int(rng.random() < p)``
    on every node of the lattice.  Decoding such a realisation needs uniform weights
    (``weight_options={"method": "uniform"}``, the reference's default).

    One documented difference: the draws are consumed in node-index order (sorted coordinates), as
    one ``rng.random(N)`` call; the reference walks ``egraph.nodes`` in the insertion order its
    ``RHG_graph`` happens to produce from a Python ``set`` (``surface_code.py:186-226``), which
    is an artefact of tuple hashing, not a specification.  Same distribution; to replay a reference
    realisation inject its bits (``TrialEngine.run_bits`` or ``IidNoise.set_bits``).
    """

    def __init__(self, code, error_probability):
        if error_probability < 0.0 or error_probability > 1.0:
            raise ValueError("Probability is not between 0 and 1.")
        self.code = code
        self.egraph = code.graph
        self.error_probability = float(error_probability)

    def set_bits(self, bit_val):
        """Install a given realisation: ``bit_val`` for all N nodes in node-index order."""
        bits = (np.asarray(bit_val).reshape(-1) != 0).astype(np.int8)
        if bits.shape[0] != self.code.lattice.N:
            raise ValueError(f"need one bit per node ({self.code.lattice.N})")
        g = self.egraph
        g._ensure_arrays()
        g.bit_val[:] = bits
        self.code._trial = ("bits", bits.astype(np.uint8))

    def apply_noise(self, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        self.set_bits(rng.random(self.code.lattice.N) < self.error_probability)
