"""`LapTimeEvaluator` -- the batched entry point: alphas[B, N_alpha] -> lap[B] on one B200.

Owns one `ltk_ctx` (track constants, vehicle constants, sampling), a reusable device workspace and a
pinned host staging buffer.  Inputs may be torch CUDA tensors (stay resident) or host arrays (copied
through pinned memory).  Everything numeric happens in libltk's kernels."""
from __future__ import annotations

import ctypes as C
import os
import sys
import time
import math

import numpy as np

from . import _device, _native

DEFAULT_TOPK = 10  # trajectory_bayesian_nonlinear.py:257


class Lane:
    """One population in flight: an evaluator (context + workspace) and the stream its kernels run on."""

    def __init__(self, ev, stream):
        self.ev, self.stream = ev, stream


class LapTimeEvaluator:
    SPLINE_MODES = {"tridiagonal": 0, "fitpack": 1}  # LTK_SPLINE_* (include/ltk.h)

    def __init__(self, track, vehicle, mode="bayes", ns=None, device=None, max_workspace_bytes=None,
                 spline=None):
        torch = _device.torch_cuda()
        if spline is None:
            from .path import default_spline

            spline = default_spline()
        self.torch = torch
        self.lib = _native.load()
        self.track, self.vehicle, self.mode = track, vehicle, mode
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.ns = int(math.ceil(track.length) if ns is None else ns)  # trajectory.py:35
        left, diff = track.affine_map(mode)
        self.n_alpha = left.shape[1]
        self._veh = vehicle.to_ltk()
        handle = C.c_void_p()
        dp = C.POINTER(C.c_double)
        rc = self.lib.ltk_create(C.byref(handle), self.device.index, left.ctypes.data_as(dp),
                                 diff.ctypes.data_as(dp), self.n_alpha, C.byref(self._veh), self.ns)
        _native.check(rc)
        self._ctx = handle
        self._ws = None
        self._pinned = None
        self._pinned_out = None
        if max_workspace_bytes is None:
            free, _total = torch.cuda.mem_get_info(self.device)
            max_workspace_bytes = int(free * 0.8)
        self.max_workspace_bytes = max_workspace_bytes
        self.spline = "tridiagonal"
        self._split_on = True
        if spline != "tridiagonal":
            self.set_spline_mode(spline)

    def set_spline_mode(self, spline):
        """"tridiagonal" (default, fastest): classical cyclic-tridiagonal periodic spline.  "fitpack": SciPy
        FITPACK's own arithmetic (fpclos Givens QR, splder) -- the bits `splprep(per=1)` / `splev` give the
        reference (path.py:25, :51-54); lap times then match the reference bit for bit on most candidates."""
        _native.check(self.lib.ltk_set_spline_mode(self._ctx, self.SPLINE_MODES[spline]), self._ctx)
        self.spline = spline
        self._ws = None  # the workspace layout depends on the mode
        for lane in (getattr(self, "_lanes", None) or [])[1:]:
            lane.ev.set_spline_mode(spline)

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        for lane in (getattr(self, "_lanes", None) or [])[1:]:
            lane.ev.close()
        self._lanes = None
        if getattr(self, "_ctx", None):
            self.lib.ltk_destroy(self._ctx)
            self._ctx = None
        self._ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_ns(self, ns):
        """`Trajectory.ns` is a plain attribute in the reference; changing it re-samples the lap."""
        _native.check(self.lib.ltk_set_ns(self._ctx, int(ns)), self._ctx)
        self.ns = int(ns)
        for lane in (getattr(self, "_lanes", None) or [])[1:]:
            lane.ev.set_ns(ns)

    # -- tracing ------------------------------------------------------------------------------------
    def trace_begin(self, max_records=256):
        """Record a CUDA event before and after every pipeline kernel of this evaluator and of its lanes."""
        for ev in [self] + [lane.ev for lane in (getattr(self, "_lanes", None) or [])[1:]]:
            _native.check(self.lib.ltk_trace_begin(ev._ctx, int(max_records)), ev._ctx)

    def trace_read(self):
        """[(lane, kernel, start_ms, end_ms)] of everything recorded since `trace_begin`, on one time line
        (zero = the first kernel this evaluator launched), sorted by start; switches tracing off."""
        names = {1: "k1a", 2: "k1b", 3: "k23"}
        evs = [self] + [lane.ev for lane in (getattr(self, "_lanes", None) or [])[1:]]
        out = []
        for li, ev in enumerate(evs):
            cap = 4096
            kind, t0, t1, n = (C.c_int * cap)(), (C.c_float * cap)(), (C.c_float * cap)(), C.c_int()
            _native.check(self.lib.ltk_trace_read(ev._ctx, self._ctx, cap, kind, t0, t1, C.byref(n)), ev._ctx)
            out += [(li, names.get(kind[i], "?"), float(t0[i]), float(t1[i])) for i in range(n.value)]
        for ev in evs:
            _native.check(self.lib.ltk_trace_begin(ev._ctx, 0), ev._ctx)
        return sorted(out, key=lambda r: r[2])

    def set_sweep_precision(self, bits):
        """64 (default): everything in fp64.  32: the optional fp32 variant of the velocity sweeps (spline and
        curvature stay fp64; lap times then agree with the fp64 ones to ~1e-5, tolerance 1e-4)."""
        _native.check(self.lib.ltk_set_sweep_precision(self._ctx, int(bits)), self._ctx)
        for lane in (getattr(self, "_lanes", None) or [])[1:]:
            lane.ev.set_sweep_precision(bits)
        self.sweep_bits = int(bits)

    # -- lanes: several populations in flight ---------------------------------------------------------
    def lanes(self, n=3):
        """`n` independent (evaluator, stream) pairs for scoring several populations CONCURRENTLY.

        The kernels of one population leave the GPU partly idle -- the spline solve is latency-bound at a
        few warps per SM, every kernel has a tail -- so populations issued on different streams overlap:
        measured on B200, 65,536 Buckmore/TBR18 candidates: 0.97 ms per population on one stream, 0.83 ms
        with three in flight.  Each lane owns a context (its top-k scratch is per context) and a
        workspace; lane 0 is this evaluator on its own stream."""
        torch = self.torch
        cur = getattr(self, "_lanes", None) or []
        if n > 1 and len(cur) < n:  # the lanes share the memory budget
            self.max_workspace_bytes = min(self.max_workspace_bytes, int(torch.cuda.mem_get_info(self.device)[0] * 0.8) // n)
        while len(cur) < n:
            ev = self if not cur else LapTimeEvaluator(self.track, self.vehicle, self.mode, self.ns, self.device.index,
                                                       self.max_workspace_bytes, spline=self.spline)
            if cur:
                ev.wave_lanes = 1  # a lane scores its population itself; only the owner fans out
                if getattr(self, "sweep_bits", 64) != 64:
                    ev.set_sweep_precision(self.sweep_bits)
            cur.append(Lane(ev, torch.cuda.Stream(self.device)))
        self._lanes = cur
        # several populations in flight: one sweep launch each (see ltk_set_sweep_split); a later
        # single-population call on this evaluator switches the split back on (lap_times_device)
        for lane in cur[:n]:
            lane.ev._set_split(n <= 1)
        return cur[:n]

    def _set_split(self, on):
        if self._split_on != bool(on):
            _native.check(self.lib.ltk_set_sweep_split(self._ctx, int(bool(on))), self._ctx)
            self._split_on = bool(on)

    # -- sizing -----------------------------------------------------------------------------------
    def workspace_bytes(self, B):
        out = C.c_size_t()
        _native.check(self.lib.ltk_workspace_bytes(self._ctx, int(B), C.byref(out)), self._ctx)
        return out.value

    def max_batch(self):
        """Largest candidate count whose workspace fits the budget (multiple of 4096)."""
        per = self.workspace_bytes(4096) / 4096.0
        b = int(self.max_workspace_bytes / per) // 4096 * 4096
        return max(b, 32)

    def _workspace(self, B):
        need = self.workspace_bytes(B)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
        # lanes run on streams other than the one the block was allocated on
        self._ws.record_stream(self.torch.cuda.current_stream(self.device))
        return self._ws

    # -- evaluation -------------------------------------------------------------------------------
    def lap_times_device(self, alphas, out=None, _lane=False):
        """alphas: float64 CUDA tensor [B, n_alpha] (contiguous) -> float64 CUDA tensor [B].
        Asynchronous on the current stream.  (`_lane`: set by the multi-lane drivers, which manage the
        sweep split themselves.)"""
        torch = self.torch
        if alphas.dtype != torch.float64 or not alphas.is_cuda:
            raise ValueError("alphas must be a float64 CUDA tensor")
        alphas = alphas.contiguous()
        if alphas.dim() != 2 or alphas.shape[1] != self.n_alpha:
            raise ValueError(f"alphas must be [B, {self.n_alpha}]")
        B = alphas.shape[0]
        if out is None:
            out = torch.empty(B, dtype=torch.float64, device=self.device)
        if B >= 2 * self.WAVE and self.wave_lanes > 1:
            return self._lap_times_waves(alphas, out)
        if not _lane:
            self._set_split(True)
        chunk = self.max_batch()
        st = _device.stream_ptr(torch, self.device)
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            ws = self._workspace(hi - lo)
            rc = self.lib.ltk_eval_alphas(self._ctx, _device.ptr(alphas[lo:hi]), hi - lo, _device.ptr(out[lo:hi]),
                                          _device.ptr(ws), ws.numel(), st)
            _native.check(rc, self._ctx)
        return out

    # A population much larger than one resident wave is scored as a sequence of single-wave chunks spread
    # over the lanes: measured (B200, 2^20 Buckmore/TBR18 candidates) 17.2 ms as one multi-wave launch
    # sequence with a 15 GB workspace, against ~13 ms in 65,536-candidate chunks three at a time -- the
    # chunks overlap each other's latency-bound phases and the workspace stays at 3 x 0.96 GB.
    # `lap_times` batches up to here go through ltk_eval_alphas_host (measured: 247 / 256 / 285 us per call at
    # 1 / 44 / 1,024 candidates against 284 / 289 / 302 us through the torch-staged route; from 8,192 on the
    # single-threaded copy into the staging buffer makes it the slower one)
    HOST_GRAPH_MAX = 2048
    WAVE = 65536  # measured against 75,776 (= 4 warps on every scheduler): 12.9 vs 13.6 ms for 2^20 candidates
    wave_lanes = 3

    def _lap_times_waves(self, alphas, out):
        torch = self.torch
        B = alphas.shape[0]
        pool = self.lanes(self.wave_lanes)
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        chunk = min(self.WAVE, self.max_batch())
        for i, lo in enumerate(range(0, B, chunk)):
            hi = min(B, lo + chunk)
            lane = pool[i % len(pool)]
            if i < len(pool):
                lane.stream.wait_event(ready)
            with torch.cuda.stream(lane.stream):
                ev = lane.ev
                ws = ev._workspace(hi - lo)
                rc = ev.lib.ltk_eval_alphas(ev._ctx, _device.ptr(alphas[lo:hi]), hi - lo, _device.ptr(out[lo:hi]),
                                            _device.ptr(ws), ws.numel(), _device.stream_ptr(torch, self.device))
                _native.check(rc, ev._ctx)
        for lane in pool:
            done = torch.cuda.Event()
            done.record(lane.stream)
            main.wait_event(done)
        alphas.record_stream(pool[0].stream)
        return out

    def kernel_times(self, alphas, out, reps=5):
        """Average CUDA-event duration (ms) of each pipeline kernel over `reps` passes (measurement hook)."""
        B = alphas.shape[0]
        ws = self._workspace(B)
        ms = (C.c_float * 4)()
        acc = np.zeros(4)
        for _ in range(reps):
            rc = self.lib.ltk_eval_alphas_timed(self._ctx, _device.ptr(alphas), B, _device.ptr(out), _device.ptr(ws),
                                                ws.numel(), _device.stream_ptr(self.torch, self.device), ms)
            _native.check(rc, self._ctx)
            acc += np.array(ms[:])
        acc /= reps
        return {"k1a_spline_solve": float(acc[0]), "k1b_curvature": float(acc[1]), "k23_sweep": float(acc[2])}

    def random_population_device(self, count, key, first_row=0, low=0.0, high=0.99):
        """`count` candidates with every alpha ~ U[low, high) generated ON THE DEVICE (tbn.py:142, :244 draw them
        one element at a time with np.random.uniform).  `key` = (k0, k1) selects the stream; rows
        [first_row, first_row + count) of the population
            np.random.Generator(np.random.Philox(key=key)).uniform(low, high, (total_rows, n_alpha))
        come out bit for bit, so ranks take disjoint row ranges of one reproducible population."""
        torch = self.torch
        out = torch.empty((int(count), self.n_alpha), dtype=torch.float64, device=self.device)
        rc = self.lib.ltk_random_uniform(self.device.index, int(key[0]), int(key[1]), int(first_row) * self.n_alpha,
                                         out.numel(), float(low), float(high), _device.ptr(out),
                                         _device.stream_ptr(torch, self.device))
        _native.check(rc)
        return out

    def curvature_objectives_device(self, alphas):
        """alphas: float64 CUDA tensor [B, n_alpha] -> (gamma2[B], length[B]) CUDA tensors: the sum of squared
        sample curvatures `Path.gamma2(s)` and `Path.length` of every candidate's spline (the objectives of
        `Trajectory.minimise_curvature` / `minimise_compromise`, trajectory.py:60-97)."""
        torch = self.torch
        if alphas.dtype != torch.float64 or not alphas.is_cuda:
            raise ValueError("alphas must be a float64 CUDA tensor")
        alphas = alphas.contiguous()
        if alphas.dim() != 2 or alphas.shape[1] != self.n_alpha:
            raise ValueError(f"alphas must be [B, {self.n_alpha}]")
        B = alphas.shape[0]
        g2 = torch.empty(B, dtype=torch.float64, device=self.device)
        length = torch.empty(B, dtype=torch.float64, device=self.device)
        chunk = self.max_batch()
        st = _device.stream_ptr(torch, self.device)
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            ws = self._workspace(hi - lo)
            rc = self.lib.ltk_eval_objectives(self._ctx, _device.ptr(alphas[lo:hi]), hi - lo, _device.ptr(g2[lo:hi]),
                                              _device.ptr(length[lo:hi]), _device.ptr(ws), ws.numel(), st)
            _native.check(rc, self._ctx)
        return g2, length

    def curvature_objectives(self, alphas):
        """Host in, host out version of `curvature_objectives_device`."""
        a = np.ascontiguousarray(alphas, dtype=np.float64)
        if a.ndim == 1:
            a = a.reshape(1, -1)
        g2, length = self.curvature_objectives_device(_device.to_device(a, self.device))
        return g2.cpu().numpy(), length.cpu().numpy()

    def merge_topk_device(self, laps, idx, k=DEFAULT_TOPK):
        """Stable ascending top-k of explicit (lap, global index) pairs (the multi-GPU merge)."""
        torch = self.torch
        best = torch.empty(k, dtype=torch.float64, device=self.device)
        out_idx = torch.empty(k, dtype=torch.int64, device=self.device)
        rc = self.lib.ltk_topk_pairs(self._ctx, _device.ptr(laps), _device.ptr(idx), laps.numel(), int(k),
                                     _device.ptr(best), _device.ptr(out_idx), _device.stream_ptr(torch, self.device))
        _native.check(rc, self._ctx)
        return best, out_idx

    def merge_gathered_device(self, gathered, world, k_in, k=DEFAULT_TOPK):
        """Stable top-k of an all-gathered buffer [world][2][k_in] (int64: lap bit patterns, then global indices per
        rank -- `lap_times_topk_device(..., packed=True)`'s fourth result on every rank) in one launch."""
        torch = self.torch
        best = torch.empty(k, dtype=torch.float64, device=self.device)
        out_idx = torch.empty(k, dtype=torch.int64, device=self.device)
        rc = self.lib.ltk_topk_gathered(self._ctx, _device.ptr(gathered), int(world), int(k_in), int(k),
                                        _device.ptr(best), _device.ptr(out_idx), _device.stream_ptr(torch, self.device))
        _native.check(rc, self._ctx)
        return best, out_idx

    def controls_lap_times_device(self, xy, out=None):
        """xy: float64 CUDA tensor [B, 2, n_alpha + 1] of control points (calcMinTime surface)."""
        torch = self.torch
        if xy.dtype != torch.float64 or not xy.is_cuda:
            raise ValueError("controls must be a float64 CUDA tensor")
        xy = xy.contiguous()
        if xy.dim() != 3 or xy.shape[1] != 2 or xy.shape[2] != self.n_alpha + 1:
            raise ValueError(f"controls must be [B, 2, {self.n_alpha + 1}]")
        B = xy.shape[0]
        if out is None:
            out = torch.empty(B, dtype=torch.float64, device=self.device)
        chunk = self.max_batch()
        st = _device.stream_ptr(torch, self.device)
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            ws = self._workspace(hi - lo)
            rc = self.lib.ltk_eval_controls(self._ctx, _device.ptr(xy[lo:hi]), xy.shape[2], hi - lo,
                                            _device.ptr(out[lo:hi]), _device.ptr(ws), ws.numel(), st)
            _native.check(rc, self._ctx)
        return out

    def lap_times(self, alphas):
        """Host in, host out: numpy [B, n_alpha] -> numpy [B]."""
        a = np.ascontiguousarray(alphas, dtype=np.float64)
        if a.ndim == 1:
            a = a.reshape(1, -1)
        B = a.shape[0]
        if a.shape[1] != self.n_alpha:
            raise ValueError(f"alphas must be [B, {self.n_alpha}]")
        if 0 < B <= self.HOST_GRAPH_MAX:
            # the optimiser loops' latency path: one native call, one CUDA graph per batch size
            out = np.empty(B, dtype=np.float64)
            _native.check(self.lib.ltk_eval_alphas_host(self._ctx, a.ctypes.data, B, out.ctypes.data), self._ctx)
            return out
        return self._lap_times_staged(a)

    def _lap_times_staged(self, a):
        """The torch route of `lap_times`: pinned staging tensors, the current stream, any batch size."""
        torch = self.torch
        B = a.shape[0]
        if self._pinned is None or self._pinned.shape[0] < B:
            self._pinned = torch.empty((B, self.n_alpha), dtype=torch.float64).pin_memory()
            self._pinned_out = torch.empty(B, dtype=torch.float64).pin_memory()
        self._pinned[:B].copy_(torch.from_numpy(a))
        d_a = self._pinned[:B].to(self.device, non_blocking=True)
        d_lap = self.lap_times_device(d_a)
        self._pinned_out[:B].copy_(d_lap, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._pinned_out[:B].numpy().copy()

    def stream_populations(self, populations, k=DEFAULT_TOPK, index_base=0, index_stride=None, finish=None, lanes=3,
                           slots=None):
        """Score a SEQUENCE of host populations with copies and kernels overlapped.

        `populations` yields float64 arrays [B, n_alpha]: pinned torch tensors are copied as they are,
        numpy arrays / pageable tensors go through a pinned staging buffer first.  `lanes` populations
        are in flight at a time, each on its own compute stream (see `lanes()`); one copy stream moves
        the next populations host->device and a second one the finished results device->host, so
        uploads, kernels and downloads of different populations overlap.  Input / result buffers come in
        `slots` sets (default 2 x lanes): the host runs that many populations ahead of the results it hands
        out, so a lane that finishes a population finds the next one already uploaded (with as many slots as
        lanes the upload of population i could only be issued once population i - lanes had been taken,
        and its lane idled through the copy).  Yields, in submission order, one `(laps, best_laps,
        best_idx)` triple of numpy views per population; the views alias pinned slot buffers and stay
        valid until the next result is requested.  `finish(best, idx) ->
        (best, idx)` runs on the population's compute stream after the local top-k (the multi-GPU
        all-gather + merge hooks in here).  Candidate j of population i gets the global index
        index_base + i*index_stride + j (index_stride defaults to the population size).  This is the
        path `bench.py` times as `e2e`."""
        torch = self.torch
        dev = self.device
        pool = self.lanes(max(1, int(lanes)))
        # one copy stream per direction: on a single in-order stream the upload of the next population
        # would queue behind the download of the previous one, which waits for that one's kernels
        if getattr(self, "_copy_streams", None) is None:
            self._copy_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            # LTK_E2E_COPY_STREAMS=2 (A/B): uploads alternate between two streams, two host->device copies in flight
            self._copy_in_extra = [torch.cuda.Stream(dev)
                                   for _ in range(max(0, int(os.environ.get("LTK_E2E_COPY_STREAMS", "1")) - 1))]
        copy_in, copy_out = self._copy_streams
        copy_ins = [copy_in] + self._copy_in_extra
        # join the caller's stream: lane 0 is this evaluator (its workspace, top-k scratch and ticket), and
        # earlier asynchronous calls on the current stream may still be using them
        entry = torch.cuda.Event()
        entry.record(torch.cuda.current_stream(dev))
        for st_ in [lane.stream for lane in pool] + copy_ins + [copy_out]:
            st_.wait_event(entry)
        nslot = max(len(pool), int(slots) if slots else 2 * len(pool))
        slots = [None] * nslot
        pending = []  # (slot index, B) in submission order

        def make_slot(B):
            return {"B": B,
                    "d_in": torch.empty((B, self.n_alpha), dtype=torch.float64, device=dev),
                    "d_lap": torch.empty(B, dtype=torch.float64, device=dev),
                    "h_stage": None,
                    "h_lap": torch.empty(B, dtype=torch.float64).pin_memory(),
                    "h_best": torch.empty(k, dtype=torch.float64).pin_memory(),
                    "h_idx": torch.empty(k, dtype=torch.int64).pin_memory(),
                    "ev_in": torch.cuda.Event(), "ev_done": torch.cuda.Event(), "ev_out": torch.cuda.Event(),
                    "ev_fin": torch.cuda.Event(),
                    "used": False}

        def take(entry):
            si, B = entry
            sl = slots[si]
            sl["ev_out"].synchronize()
            return sl["h_lap"][:B].numpy(), sl["h_best"].numpy(), sl["h_idx"].numpy()

        base = int(index_base)
        # LTK_E2E_PROFILE=1: where the host thread's time goes (blocked on results / in the cross-rank hook / the rest)
        prof = {"take": 0.0, "finish": 0.0, "n": 0} if os.environ.get("LTK_E2E_PROFILE") else None
        t_loop = time.perf_counter()
        for i, pop in enumerate(populations):
            si = i % nslot
            lane = pool[i % len(pool)]
            if prof is not None:
                prof["n"] = i + 1
            if len(pending) == nslot:  # the slot about to be reused still holds an untaken result
                t0_ = time.perf_counter()
                res_ = take(pending.pop(0))
                if prof is not None:
                    prof["take"] += time.perf_counter() - t0_
                yield res_
            t = pop if hasattr(pop, "is_pinned") else torch.from_numpy(np.ascontiguousarray(pop, dtype=np.float64))
            if t.dim() != 2 or t.shape[1] != self.n_alpha or t.dtype != torch.float64:
                raise ValueError(f"populations must be float64 [B, {self.n_alpha}]")
            B = t.shape[0]
            sl = slots[si]
            if sl is None or sl["B"] < B or sl["h_best"].numel() != k:
                sl = slots[si] = make_slot(B)
            if not t.is_pinned():
                if sl["h_stage"] is None or sl["h_stage"].shape[0] < B:
                    sl["h_stage"] = torch.empty((B, self.n_alpha), dtype=torch.float64).pin_memory()
                if sl["used"]:
                    sl["ev_in"].synchronize()  # the previous H2D out of this staging buffer
                sl["h_stage"][:B].copy_(t)
                t = sl["h_stage"][:B]
            cin = copy_ins[i % len(copy_ins)]
            with torch.cuda.stream(cin):
                if sl["used"]:
                    cin.wait_event(sl["ev_done"])  # kernels that read d_in of this slot
                sl["d_in"][:B].copy_(t, non_blocking=True)
                sl["ev_in"].record(cin)
            with torch.cuda.stream(lane.stream):
                lane.stream.wait_event(sl["ev_in"])
                if sl["used"]:
                    lane.stream.wait_event(sl["ev_out"])  # d_lap of this slot has been read back
                d_lap, best, idx, pk = lane.ev.lap_times_topk_device(sl["d_in"][:B], out=sl["d_lap"][:B], k=k,
                                                                     index_base=base, _lane=len(pool) > 1, packed=True)
                sl["ev_done"].record(lane.stream)
            if finish is not None:
                # the cross-rank step gets an event of its own: the slot's upload buffer is free again once this
                # rank's kernels are done (ev_done), and the lap times go home then -- only the k best wait for the
                # other ranks (with one event for both, every rank's uploads stalled behind the slowest rank's)
                fin = sl["ev_done"] if os.environ.get("LTK_E2E_COUPLED") else sl["ev_fin"]  # (A/B: the old single event)
                t0_ = time.perf_counter()
                best, idx = self._finish_async(finish, best, idx, sl["ev_done"], fin, pk)
                if prof is not None:
                    prof["finish"] += time.perf_counter() - t0_
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(sl["ev_done"])
                sl["h_lap"][:B].copy_(d_lap, non_blocking=True)
                if finish is not None:
                    copy_out.wait_event(fin)
                sl["h_best"].copy_(best, non_blocking=True)
                sl["h_idx"].copy_(idx, non_blocking=True)
                sl["ev_out"].record(copy_out)
            sl["keep"] = (best, idx)  # read by the copy stream after this generator moves on
            sl["used"] = True
            pending.append((si, B))
            base += B if index_stride is None else int(index_stride)
        if prof is not None:
            n_ = max(1, prof["n"])
            print(f"[stream_populations] host per population: loop {1e6 * (time.perf_counter() - t_loop) / n_:.0f} us, of which blocked in take "
                  f"{1e6 * prof['take'] / n_:.0f} us, in finish {1e6 * prof['finish'] / n_:.0f} us", file=sys.stderr, flush=True)
        while pending:
            yield take(pending.pop(0))

    def run_resident(self, populations, outs, k=DEFAULT_TOPK, index_base=0, finish=None, lanes=3):
        """Score device-resident populations `lanes` at a time (see `lanes()`): populations[i] -> outs[i % lanes]
        (float64 CUDA tensors), local top-k (+ `finish`) after each.  Enqueues only; the calling stream is
        joined to every lane before returning, so CUDA events recorded around the call time all of it.
        Returns the last (best, idx)."""
        torch = self.torch
        pool = self.lanes(max(1, int(lanes)))
        main = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(main)
        last = None
        for lane in pool:
            lane.stream.wait_event(start)
        for i, pop in enumerate(populations):
            lane = pool[i % len(pool)]
            with torch.cuda.stream(lane.stream):
                res = lane.ev.lap_times_topk_device(pop, out=outs[i % len(pool)], k=k, index_base=index_base,
                                                    _lane=len(pool) > 1, packed=True)
                last, pk = res[1:3], res[3]
                if finish is not None:
                    done = torch.cuda.Event()
                    done.record(lane.stream)
            if finish is not None:
                last = self._finish_async(finish, last[0], last[1], done, None, pk)
        for st in [lane.stream for lane in pool] + ([self._comm_stream] if finish is not None else []):
            done = torch.cuda.Event()
            done.record(st)
            main.wait_event(done)
        return last

    def _finish_async(self, finish, best, idx, ready, done=None, packed=None):
        """Run the cross-rank step (all-gather + merge) of one population on a communication stream of its
        own: the lane that produced (best, idx) goes straight on to its next population instead of
        waiting out the collective's latency.  `done` (default: `ready` itself) is recorded when the result is
        complete."""
        torch = self.torch
        if getattr(self, "_comm_stream", None) is None:
            self._comm_stream = torch.cuda.Stream(self.device)
        comm = self._comm_stream
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            best.record_stream(comm)
            idx.record_stream(comm)
            if packed is not None and getattr(finish, "packed", False):
                packed.record_stream(comm)
                out = finish(packed)  # one collective on the packed list, one merge launch (distributed.PackedTopkGather)
            else:
                out = finish(best, idx)
            (done if done is not None else ready).record(comm)
        return out

    def lap_times_topk_device(self, alphas, out=None, k=DEFAULT_TOPK, index_base=0, _lane=False, packed=False):
        """`lap_times_device` and `topk_device` of the same population in one call: -> (laps[B], best[k], idx[k])
        CUDA tensors.  A population that fits one chunk goes through `ltk_eval_alphas_topk` -- the k best are
        selected in the sweep kernel's epilogue, no separate selection launch; anything else (multi-wave
        populations, chunked workspaces) scores first and selects afterwards.  Asynchronous on the current stream.
        `packed=True` adds a fourth result: the int64 tensor [2k] that `best` (as bit patterns) and `idx` are the two
        halves of -- the unit of the multi-GPU all-gather (`distributed.PackedTopkGather`)."""
        torch = self.torch
        B = alphas.shape[0]
        if (alphas.dtype != torch.float64 or not alphas.is_cuda or alphas.dim() != 2 or alphas.shape[1] != self.n_alpha
                or not alphas.is_contiguous() or B < 1 or (B >= 2 * self.WAVE and self.wave_lanes > 1)
                or B > self.max_batch()):
            laps = self.lap_times_device(alphas, out=out, _lane=_lane)
            best, idx = self.topk_device(laps, k, index_base=index_base)
            if packed:
                pk = torch.cat([best.view(torch.int64), idx])
                return laps, pk[:k].view(torch.float64), pk[k:], pk
            return laps, best, idx
        if out is None:
            out = torch.empty(B, dtype=torch.float64, device=self.device)
        if not _lane:
            self._set_split(True)
        pk = torch.empty(2 * k, dtype=torch.int64, device=self.device)
        best, idx = pk[:k].view(torch.float64), pk[k:]
        ws = self._workspace(B)
        rc = self.lib.ltk_eval_alphas_topk(self._ctx, _device.ptr(alphas), B, _device.ptr(out), _device.ptr(ws),
                                           ws.numel(), int(index_base), int(k), _device.ptr(best), _device.ptr(idx),
                                           _device.stream_ptr(torch, self.device))
        _native.check(rc, self._ctx)
        return (out, best, idx, pk) if packed else (out, best, idx)

    def topk_device(self, laps, k=DEFAULT_TOPK, index_base=0):
        """Stable ascending top-k of a CUDA lap tensor -> (lap[k], idx[k]) CUDA tensors."""
        torch = self.torch
        best = torch.empty(k, dtype=torch.float64, device=self.device)
        idx = torch.empty(k, dtype=torch.int64, device=self.device)
        rc = self.lib.ltk_topk(self._ctx, _device.ptr(laps), laps.numel(), int(index_base), int(k),
                               _device.ptr(best), _device.ptr(idx), _device.stream_ptr(torch, self.device))
        _native.check(rc, self._ctx)
        return best, idx

    def topk(self, laps, k=DEFAULT_TOPK, index_base=0):
        if isinstance(laps, np.ndarray):
            laps = _device.to_device(laps, self.device)
        best, idx = self.topk_device(laps, k, index_base)
        return best.cpu().numpy(), idx.cpu().numpy()

    def profile(self, alpha):
        """Everything the reference exposes for ONE candidate (natural sample order), as numpy."""
        torch = self.torch
        a = _device.to_device(np.asarray(alpha, dtype=np.float64).reshape(-1), self.device)
        if a.numel() != self.n_alpha:
            raise ValueError(f"alpha must have {self.n_alpha} entries")
        n = self.ns - 1
        names = ("k", "v_local", "v_acclim", "v_declim", "v")
        d_s = torch.empty(self.ns, dtype=torch.float64, device=self.device)
        bufs = {nm: torch.empty(n, dtype=torch.float64, device=self.device) for nm in names}
        scal = torch.empty(2, dtype=torch.float64, device=self.device)
        rc = self.lib.ltk_profile(self._ctx, _device.ptr(a), _device.ptr(d_s), *[_device.ptr(bufs[nm]) for nm in names],
                                  _device.ptr(scal), _device.stream_ptr(torch, self.device))
        _native.check(rc, self._ctx)
        out = {nm: b.cpu().numpy() for nm, b in bufs.items()}
        out["s"] = d_s.cpu().numpy()
        sc = scal.cpu().numpy()
        out["lap"], out["length"] = np.float64(sc[0]), np.float64(sc[1])
        return out
