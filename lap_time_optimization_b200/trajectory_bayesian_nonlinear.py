"""`TrajectoryBayesianNonlinear` facade (reference src/trajectory_bayesian_nonlinear.py:22-80, :230-257).

Keeps `update`, `lap_time`, `updateAlphas`, `calcMinTime` and adds the batched population stage the
reference runs one candidate at a time (`Nonlinear`, :239-257; `Bayesian` database, :136-160):
`lap_time_batch`, `random_population`, `population_topk`."""
from __future__ import annotations

import math
import time

import numpy as np

from . import _device
from .evaluator import DEFAULT_TOPK, LapTimeEvaluator
from .path import Path
from .velocity import VelocityProfile

ALPHA_LOW, ALPHA_HIGH = 0.0, 0.99  # trajectory_bayesian_nonlinear.py:142, :244


class TrajectoryBayesianNonlinear:
    MODE = "bayes"

    def __init__(self, track, vehicle, device=None):
        self.track = track
        self.ns = math.ceil(track.length)
        self.vehicle = vehicle
        self._device = device
        self._evaluator = None
        self._velocity = None
        self._lap = None
        self.update(np.full(track.size, 0.5))
        self.direction_vector = track.diffs
        self.length = track.length
        self.mid_waipoints = track.mid_controls
        self.mid_controls_decongested = track.mid_controls_decongested
        self.widths_decongested = track.widths_decongested
        self.best = []
        self.sigma = []

    @property
    def evaluator(self) -> LapTimeEvaluator:
        if self._evaluator is None:
            self._evaluator = LapTimeEvaluator(self.track, self.vehicle, self.MODE, self.ns, self._device)
        elif self._evaluator.ns != self.ns:
            self._evaluator.set_ns(self.ns)
        return self._evaluator

    @property
    def n_alpha(self):
        return self.track.left_decongested.shape[1] - 1

    # -- reference surface ------------------------------------------------------------------------
    def update(self, alphas):
        """Full-resolution control points and path (tbn.py:44-49)."""
        self.alphas = alphas
        self.path = Path(self.track.control_points(alphas), self.track.closed)
        self.s = np.linspace(0, self.path.length, self.ns)

    @property
    def velocity(self):
        """Velocity profile of the last `calcMinTime` path, materialised on first access."""
        if self._velocity is None and self._lap is not None:
            s = self.s[:-1]
            s_max = self.path.length if self.track.closed else None
            self._velocity = VelocityProfile(self.vehicle, s, self.path.curvature(s), s_max)
        return self._velocity

    @velocity.setter
    def velocity(self, value):
        self._velocity = value

    def lap_time(self):
        """Lap time of the last evaluated path (tbn.py:51-54)."""
        return self._lap

    def updateAlphas(self, alphas):
        """alphas on the every-3rd-cone subset -> closed control polygon [2, m] (tbn.py:58-62)."""
        return Path(self.track.control_points_bayesian(alphas), self.track.closed).controls

    def calcMinTime(self, controls):
        """Minimum lap time along the spline through `controls` (tbn.py:65-80)."""
        controls = np.asarray(controls, dtype=np.float64)
        self.path = Path(controls, self.track.closed)
        self.s = np.linspace(0, self.path.length, self.ns)
        self._velocity = None
        ev = self.evaluator
        xy = _device.to_device(controls.reshape(1, 2, -1), ev.device)
        self._lap = np.float64(ev.controls_lap_times_device(xy).cpu().numpy()[0])
        return self._lap

    # -- batched surface --------------------------------------------------------------------------
    def lap_time_batch(self, alphas):
        """alphas [B, n_alpha] (numpy or CUDA tensor) -> lap times [B] of the same kind."""
        if isinstance(alphas, np.ndarray) or not hasattr(alphas, "is_cuda"):
            return self.evaluator.lap_times(alphas)
        return self.evaluator.lap_times_device(alphas)

    def random_population(self, count, seed=None):
        """`count` candidates with every alpha ~ U[0, 0.99) (tbn.py:142, :244)."""
        return np.random.default_rng(seed).uniform(ALPHA_LOW, ALPHA_HIGH, (count, self.n_alpha))

    def random_population_device(self, count, key, first_row=0):
        """The same population drawn on the GPU (no host->device copy): rows [first_row, first_row + count) of
        `np.random.Generator(np.random.Philox(key=key)).uniform(0, 0.99, (rows, n_alpha))`, as a CUDA tensor."""
        return self.evaluator.random_population_device(count, key, first_row, ALPHA_LOW, ALPHA_HIGH)

    def database(self, count, key, k=DEFAULT_TOPK):
        """The Bayesian method's training set (tbn.py:136-160: `X_train`, `y_train`) at population scale:
        `count` device-generated candidates and their lap times, plus the k best.  Returns CUDA tensors
        (alphas[count, n_alpha], laps[count], best_laps[k], best_idx[k])."""
        ev = self.evaluator
        d_a = self.random_population_device(count, key)
        d_lap, best, idx = ev.lap_times_topk_device(d_a, k=k)
        return d_a, d_lap, best, idx

    def population_topk(self, alphas, k=DEFAULT_TOPK):
        """Score a population and keep the k fastest: what `sorted(results)[0:10]` feeds to COBYLA
        (tbn.py:253-257).  Returns (laps[B], best_laps[k], best_indices[k]) as numpy arrays."""
        ev = self.evaluator
        d_a = alphas if hasattr(alphas, "is_cuda") else _device.to_device(alphas, ev.device)
        d_lap, best, idx = ev.lap_times_topk_device(d_a, k=k)
        return d_lap.cpu().numpy(), best.cpu().numpy(), idx.cpu().numpy()

    # -- the --nonlinear stage: random population -> k best -> COBYLA from each (tbn.py:207-270) -------
    COBYLA_MAXITER = 2000  # tbn.py:219

    def _improvement(self, tau0, tau):
        return -max(0.0, tau0 - tau)  # tbn.py:211-216

    def optimize_COBYLA(self, args, maxiter=None):
        """One COBYLA run from (tau0, alpha0) on the reference's objective (tbn.py:207-227): every
        evaluation is one candidate through the CUDA pipeline.  Returns (tau_star, w_star)."""
        from scipy.optimize import minimize

        tau0, alpha0 = args
        bounds = np.array([[ALPHA_LOW, ALPHA_HIGH] for _ in alpha0])
        ev = self.evaluator
        res = minimize(lambda x: self._improvement(tau0, ev.lap_times(x)[0]), x0=np.asarray(alpha0, dtype=np.float64),
                       bounds=bounds, method="COBYLA", options={"maxiter": maxiter or self.COBYLA_MAXITER, "disp": False})
        return float(ev.lap_times(res.x)[0]), res.x

    def optimize_COBYLA_lockstep(self, starts, maxiter=None):
        """All COBYLA runs of `Nonlinear()` at once.  The reference maps them over `Pool(processes=1)`, i.e. one
        after the other (tbn.py:256-260), and most of a run is the optimiser's own host work.  Here every
        start gets a worker PROCESS that runs scipy's COBYLA and asks this process for each objective
        value; the requests of all live workers are answered together by ONE batched pipeline call per
        round.  The workers are plain subprocesses (`_cobyla_worker.py`: no CUDA, no torch, no fork from
        this multi-threaded process, no dependence on the caller's `__main__`); results are identical to
        running `optimize_COBYLA` on each start, because a candidate's lap time does not depend on its batch."""
        import os
        import pickle
        import subprocess
        import sys

        ev = self.evaluator
        maxiter = maxiter or self.COBYLA_MAXITER
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        workers = []
        for tau0, alpha0 in starts:
            p = subprocess.Popen([sys.executable, "-m", "lap_time_optimization_b200._cobyla_worker"], stdin=subprocess.PIPE,
                                 stdout=subprocess.PIPE, env=env)
            pickle.dump((float(tau0), np.asarray(alpha0, dtype=np.float64), int(maxiter)), p.stdin)
            p.stdin.flush()
            workers.append({"proc": p, "result": None})
        live = list(range(len(workers)))
        try:
            while live:
                asks, who = [], []
                for i in list(live):
                    kind, payload = pickle.load(workers[i]["proc"].stdout)
                    if kind == "done":
                        workers[i]["result"] = payload
                        live.remove(i)
                    else:
                        asks.append(payload)
                        who.append(i)
                if asks:
                    laps = ev.lap_times(np.vstack(asks))  # one pipeline pass for this round's requests
                    for i, lap in zip(who, laps):
                        pickle.dump(float(lap), workers[i]["proc"].stdin)
                        workers[i]["proc"].stdin.flush()
        finally:
            for w in workers:
                if w["result"] is None:
                    w["proc"].kill()
                w["proc"].stdin.close()
                w["proc"].wait()
        return [(float(ev.lap_times(w["result"])[0]), w["result"]) for w in workers]

    def Nonlinear(self, population=100, starts=10, key=None, maxiter=None):
        """Racing line by random search + local refinement (tbn.py:229-270): `population` random candidates
        scored in one pass, the `starts` fastest refined by COBYLA in lock step, best of everything kept in
        `self.best` (control points).  `key` selects a device-generated population (Philox stream,
        `random_population_device`); None draws it on the host with numpy's global generator like the
        reference.  Returns the run time in seconds."""
        t0 = time.time()
        if key is None:
            alphas = np.random.uniform(ALPHA_LOW, ALPHA_HIGH, (population, self.n_alpha))
            laps, best, idx = self.population_topk(alphas, starts)
        else:
            d_a = self.random_population_device(population, key)
            laps, best, idx = self.population_topk(d_a, starts)
            alphas = d_a.cpu().numpy()
        results = list(zip(laps, alphas))
        n = min(starts, population)
        results += self.optimize_COBYLA_lockstep([(best[i], alphas[idx[i]]) for i in range(n)], maxiter)
        tau_best, alpha_best = sorted(results, key=lambda el: el[0])[0]
        self.best = self.updateAlphas(alpha_best)
        self.best_alphas, self.best_tau = np.asarray(alpha_best), float(tau_best)
        return time.time() - t0
