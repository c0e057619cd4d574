"""CPU: host-side logic (track / vehicle loaders, facade bookkeeping, C-ABI surface, multi-rank
top-k merge over gloo).  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import lap_time_optimization_b200 as ltk
from conftest import ROOT
from lap_time_optimization_b200 import _native
from lap_time_optimization_b200.distributed import allgather_topk, merge_topk, shard_bounds
from oracle import c_oracle
from oracle.reference_port import OracleTrack, load_vehicle

TRACKS = ["buckmore", "clay", "gyg", "whilton"]


@pytest.mark.parametrize("name", TRACKS)
@pytest.mark.parametrize("width", [0.8, 0.6, 1.0, 5.0, 0.0])
def test_track_matches_oracle_loader(name, width):
    tj = ltk.data_path("tracks", name + ".json")
    t, o = ltk.Track(tj, track_width=width, quiet=True), OracleTrack(tj, width)
    assert t.closed == o.closed and t.size == o.size and t.length == o.length
    for a, b in ((t.left, o.left), (t.right, o.right), (t.diffs, o.diffs), (t.left_decongested, o.left_d),
                 (t.diffs_decongested, o.diffs_d)):
        assert np.array_equal(a, b)
    rng = np.random.default_rng(5)
    a_full, a_bayes = rng.uniform(-0.5, 1.5, t.size), rng.uniform(0, 0.99, o.left_d.shape[1] - 1)
    assert np.array_equal(t.control_points(a_full), o.control_points(a_full))
    assert np.array_equal(t.control_points_bayesian(a_bayes), o.control_points_bayesian(a_bayes))
    left, diff = t.affine_map("bayes")
    assert left.shape == (2, o.left_d.shape[1] - 1) and left.flags.c_contiguous and diff.flags.c_contiguous


def test_track_sizes():
    n = {nm: ltk.Track(ltk.data_path("tracks", nm + ".json"), track_width=0.8, quiet=True) for nm in TRACKS}
    assert [n[k].left_decongested.shape[1] for k in TRACKS] == [44, 46, 40, 51]
    assert n["buckmore"].size == 131 and int(np.ceil(n["buckmore"].length)) == 847


@pytest.mark.parametrize("veh", ["tbr18.json", "MX5.json"])
def test_vehicle_constants_match_oracle(veh):
    vj = ltk.data_path("vehicles", veh)
    ours, theirs = ltk.load_vehicle(vj).to_ltk(), c_oracle.vehicle_struct(load_vehicle(vj))
    for f in ("kind", "n_map", "mass", "mu_g", "f_max", "f_max_sq", "e0", "cr2"):
        assert getattr(ours, f) == getattr(theirs, f), f
    assert list(ours.map_v) == list(theirs.map_v) and list(ours.map_f) == list(theirs.map_f)


def test_vehicle_callbacks():
    v, o = ltk.load_vehicle(ltk.data_path("vehicles", "tbr18.json")), load_vehicle(ltk.data_path("vehicles", "tbr18.json"))
    m, om = ltk.load_vehicle(ltk.data_path("vehicles", "MX5.json")), load_vehicle(ltk.data_path("vehicles", "MX5.json"))
    for vel in (0.0, 5.0, 7.3, 19.999, 35.0, 60.0):
        assert v.engine_force(vel) == o.engine_force(vel) and m.engine_force(vel) == om.engine_force(vel)
        for k in (0.0, 0.01, 0.3):
            assert v.traction(np.float64(vel), k) == o.traction(np.float64(vel), k)
            assert m.traction(np.float64(vel), k) == om.traction(np.float64(vel), k)
    assert isinstance(m, ltk.VehicleMX5) and m.friction_coef == 1.5 and v.friction_coef == 1.5


def test_path_host_attributes_and_closure_quirk():
    t = ltk.Track(ltk.data_path("tracks", "buckmore.json"), track_width=0.8, quiet=True)
    a = np.random.default_rng(0).uniform(0, 0.99, 43)
    c = t.control_points_bayesian(a)
    before = c.copy()
    p = ltk.Path(c, True)
    assert np.array_equal(p.dists, np.append(0, np.cumsum(np.linalg.norm(np.diff(before, axis=1), axis=0))))
    assert p.length == p.dists[-1]
    # like splprep(per=1): the caller's last column now equals the first
    assert np.array_equal(c[:, -1], c[:, 0]) and not np.array_equal(before[:, -1], before[:, 0])
    assert p.controls is c
    if not torch.cuda.is_available():  # Path.spline is computed by the CUDA FITPACK kernel: no CPU fallback
        with pytest.raises(ltk.LtkUnavailable):
            p.spline


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ltk.h")).read()
    declared = sorted(set(re.findall(r"\b(ltk_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 14
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libltk.so does not export {name}"
    assert sorted(_native.SIGNATURES) == declared
    assert _native.load().ltk_version() >= 100
    assert ctypes.sizeof(_native.LtkVehicle) == 8 + 8 * 4 + 2 * 8 * 16 + 16


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    t = ltk.Track(ltk.data_path("tracks", "buckmore.json"), track_width=0.8, quiet=True)
    v = ltk.load_vehicle(ltk.data_path("vehicles", "tbr18.json"))
    with pytest.raises(ltk.LtkUnavailable):
        ltk.LapTimeEvaluator(t, v)
    with pytest.raises(ltk.LtkUnavailable):
        ltk.Path(t.control_points(np.full(t.size, 0.5)), True).curvature(np.array([0.0, 1.0]))
    # the raw ABI reports an error code, it does not compute
    lib = _native.load()
    h = ctypes.c_void_p()
    left, diff = t.affine_map("bayes")
    dp = ctypes.POINTER(ctypes.c_double)
    veh = v.to_ltk()
    rc = lib.ltk_create(ctypes.byref(h), 0, left.ctypes.data_as(dp), diff.ctypes.data_as(dp), left.shape[1],
                        ctypes.byref(veh), 847)
    assert rc < 0 and lib.ltk_last_error(None)


def test_shard_bounds_cover_everything():
    for total in (0, 1, 10, 65536, 1048577):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_merge_topk_is_stable_and_handles_padding():
    laps = torch.tensor([2.0, 1.0, 1.0, float("nan"), 3.0, 0.5, 9.0], dtype=torch.float64)
    idx = torch.tensor([7, 9, 4, 1, 2, -1, 3], dtype=torch.int64)
    b, i = merge_topk(laps, idx, 4)
    assert i.tolist() == [4, 9, 7, 2] and b.tolist() == [1.0, 1.0, 2.0, 3.0]


def _gloo_worker(rank, world, port, k, q):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    laps_all = rng.uniform(40, 50, 1000)
    laps_all[17] = laps_all[900] = 39.0  # a tie across ranks: the lower global index must come first
    lo, hi = shard_bounds(1000, rank, world)
    local = torch.tensor(laps_all[lo:hi])
    order = torch.sort(local, stable=True).indices[:k]
    best, idx = allgather_topk(local[order], order + lo, k)
    q.put((rank, best.tolist(), idx.tolist()))
    dist.destroy_process_group()


def test_allgather_topk_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 10, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    rng = np.random.default_rng(11)
    laps_all = rng.uniform(40, 50, 1000)
    laps_all[17] = laps_all[900] = 39.0
    expect = sorted(range(1000), key=lambda i: laps_all[i])[:10]
    assert out[0][2] == expect and out[1][2] == expect and out[0][1] == out[1][1]
    assert expect[:2] == [17, 900]


class _HostMerge:
    """Stand-in for LapTimeEvaluator.merge_gathered_device on a machine without a GPU: the same selection from the
    gathered [world][2][k_in] layout, in torch."""

    @staticmethod
    def merge_gathered_device(gathered, world, k_in, k):
        g = gathered.view(world, 2, k_in)
        return merge_topk(g[:, 0, :].contiguous().view(torch.float64).reshape(-1), g[:, 1, :].reshape(-1), k)


def _gloo_packed_worker(rank, world, port, k, q):
    import torch.distributed as dist

    from lap_time_optimization_b200.distributed import PackedTopkGather

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(12)
    laps_all = rng.uniform(40, 50, 1001)  # uneven shards
    laps_all[3] = laps_all[777] = 38.5
    lo, hi = shard_bounds(1001, rank, world)
    local = torch.tensor(laps_all[lo:hi])
    order = torch.sort(local, stable=True).indices[:k]
    packed = torch.cat([local[order].view(torch.int64), order + lo])  # what the sweep epilogue leaves on a rank
    best, idx = PackedTopkGather(_HostMerge(), k)(packed)
    q.put((rank, best.tolist(), idx.tolist()))
    dist.destroy_process_group()


def test_packed_topk_gather_gloo_world2():
    """The multi-GPU cross-rank step (one all-gather of the packed list, one merge of the gathered layout) with two
    gloo ranks on the CPU; the merge kernel itself is checked on the GPU (test_gathered_merge_of_packed_topk_lists)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_packed_worker, args=(r, 2, port, 10, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    rng = np.random.default_rng(12)
    laps_all = rng.uniform(40, 50, 1001)
    laps_all[3] = laps_all[777] = 38.5
    expect = sorted(range(1001), key=lambda i: laps_all[i])[:10]
    assert out[0][2] == expect and out[1][2] == expect and out[0][1] == out[1][1]
    assert expect[:2] == [3, 777]


def test_result_artefact_writers(tmp_path):
    """The JSON files of src/__main__.py:196-213 / utils.py:108-136 (what mpc/track.py reads back)."""
    import json
    from types import SimpleNamespace

    from lap_time_optimization_b200 import utils

    track = ltk.Track(ltk.data_path("tracks", "buckmore.json"), track_width=0.8, quiet=True)
    s = np.linspace(0.0, 10.0, 7)
    path = SimpleNamespace(position=lambda u: np.vstack([np.cos(u), np.sin(u)]))
    traj = SimpleNamespace(path=path, s=s, velocity=SimpleNamespace(v=np.arange(6.0)))
    utils.save_result_artefacts(str(tmp_path / "plots" / "nonlinear"), track, traj)
    d = tmp_path / "plots" / "nonlinear"
    assert sorted(p.name for p in d.iterdir()) == ["left.json", "path.json", "right.json", "velocities.json", "widths.json"]
    p = json.load(open(d / "path.json"))
    assert p["name"] == "path" and p["path"]["x"] == np.cos(s).tolist() and p["path"]["y"] == np.sin(s).tolist()
    left = json.load(open(d / "left.json"))
    assert left["path"]["x"] == track.old_left[0].tolist() and len(left["path"]["y"]) == track.old_left.shape[1]
    assert json.load(open(d / "widths.json"))["width"] == track.widths.tolist()
    assert json.load(open(d / "velocities.json")) == {"name": "velocities", "velocities": [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]}


# ---- the FITPACK-mode device arithmetic, compiled for the host ------------------------------------------
@pytest.fixture(scope="module")
def fitcore(tmp_path_factory):
    """g++ build of lap_time_optimization_b200/csrc/ltk_fitpack_core.cuh (the __host__ __device__ code the
    CUDA kernels k1a_fitpack / k1b_samples<FIT> run), -ffp-contract=off like nvcc -fmad=false."""
    import subprocess

    so = str(tmp_path_factory.mktemp("fitcore") / "libfitcore.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "native", "fitcore_host.cpp"), "-lm"])
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.fitcore_solve.argtypes = [ctypes.c_int] + [dp] * 10
    lib.fitcore_curvature.argtypes = [ctypes.c_int] + [dp] * 6 + [ctypes.c_int] + [dp] * 5
    lib.fitcore_splev.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_int, dp, ctypes.c_int, dp, dp]
    lib.fitcore_solve_open.argtypes = [ctypes.c_int] + [dp] * 6
    return lib


def _fitcore_run(lib, pts, u, x):
    dp = ctypes.POINTER(ctypes.c_double)
    p = lambda a: a.ctypes.data_as(dp)  # noqa: E731
    N = pts.shape[1] - 1
    px, py = np.ascontiguousarray(pts[0, :N]), np.ascontiguousarray(pts[1, :N])
    out = {k: np.zeros(n) for k, n in (("t", N + 7), ("cx", N + 3), ("cy", N + 3), ("w1x", N + 2), ("w1y", N + 2),
                                       ("w2x", N + 1), ("w2y", N + 1))}
    u = np.ascontiguousarray(u)
    assert lib.fitcore_solve(N, p(px), p(py), p(u), *[p(out[k]) for k in ("t", "cx", "cy", "w1x", "w1y", "w2x", "w2y")]) == 0
    x = np.ascontiguousarray(x)
    ev = {k: np.zeros(x.size) for k in ("k", "dx", "dy", "ddx", "ddy")}
    lib.fitcore_curvature(N, *[p(out[k]) for k in ("t", "w1x", "w1y", "w2x", "w2y")], p(x), x.size,
                          *[p(ev[k]) for k in ("k", "dx", "dy", "ddx", "ddy")])
    out.update(ev)
    return out


def test_device_fitpack_arithmetic_equals_scipy(fitcore):
    """The single-pass register-window Givens QR of the kernels == splprep(k=3, s=0, per=1) bit for bit, and its
    derivative evaluation == splev(der=1|2), on random closed polygons of 5 .. 170 points."""
    from scipy.interpolate import splev, splprep

    rng = np.random.default_rng(21)
    for trial in range(80):
        m = int(rng.integers(6, 172))
        th = np.sort(rng.uniform(0, 2 * np.pi, m - 1))
        r = rng.uniform(50, 120, m - 1)
        pts = np.array([r * np.cos(th), r * np.sin(th)])
        pts = np.concatenate([pts, pts[:, :1]], axis=1)
        u = np.append(0, np.cumsum(np.linalg.norm(np.diff(pts, axis=1), axis=0)))
        (t, c, k), _ = splprep(pts.copy(), u=u, k=3, s=0, per=1)
        x = np.linspace(0, u[-1], 500)[:-1]
        got = _fitcore_run(fitcore, pts, u, x)
        assert np.array_equal(got["t"], t) and np.array_equal(got["cx"], c[0]) and np.array_equal(got["cy"], c[1])
        d1, d2 = splev(x, (t, c, k), der=1), splev(x, (t, c, k), der=2)
        for key, want in (("dx", d1[0]), ("dy", d1[1]), ("ddx", d2[0]), ("ddy", d2[1])):
            assert np.array_equal(got[key], want), (trial, key)


@pytest.mark.parametrize("name", ["buckmore_tbr18_bayes", "buckmore_tbr18_full", "whilton_mx5_full", "gyg_tbr18_bayes"])
def test_device_fitpack_arithmetic_equals_reference_and_oracle(fitcore, name):
    """Same code against the spline the unmodified reference built (golden tck, derivatives) and against the C
    oracle's curvature (oracle/lap_oracle.c in FITPACK mode), bit for bit."""
    from conftest import case_setup

    g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
    tj, width, vj, mode = case_setup(name)
    co = c_oracle.COracle(OracleTrack(tj, width), load_vehicle(vj), mode, int(g["ns"]), spline="fitpack")
    for i in range(int(g["n_profiles"])):
        got = _fitcore_run(fitcore, g["prof_controls"][i], g["prof_dists"][i], g["prof_s"][i][:-1])
        assert np.array_equal(got["t"], g["prof_tck_t"][i])
        assert np.array_equal(got["cx"], g["prof_tck_cx"][i]) and np.array_equal(got["cy"], g["prof_tck_cy"][i])
        for key in ("dx", "dy", "ddx", "ddy"):
            assert np.array_equal(got[key], g["prof_" + key][i]), key
        assert np.array_equal(got["k"], co.profile(g["alphas"][i])["k"])


def test_device_open_spline_and_general_evaluation_equal_scipy(fitcore):
    """The Path facade's FITPACK kernel pieces on the host: the open (not-a-knot) solve == splprep(per=0), and the
    general splev / splder evaluation (positions, first and second derivatives, end point and knots included)
    == splev, bit for bit, for open and closed splines."""
    from scipy.interpolate import splev, splprep

    dp = ctypes.POINTER(ctypes.c_double)
    p = lambda a: a.ctypes.data_as(dp)  # noqa: E731
    rng = np.random.default_rng(31)
    for trial in range(60):
        per = trial % 2
        m = int(rng.integers(6, 130))
        th = np.sort(rng.uniform(0, 1.8 * np.pi, m))
        r = rng.uniform(50, 120, m)
        pts = np.array([r * np.cos(th), r * np.sin(th)])
        if per:
            pts[:, -1] = pts[:, 0]
        u = np.append(0, np.cumsum(np.linalg.norm(np.diff(pts, axis=1), axis=0)))
        (t, c, k), _ = splprep(pts.copy(), u=u, k=3, s=0, per=per)
        if not per:
            t2, cx, cy = np.zeros(m + 4), np.zeros(m), np.zeros(m)
            assert fitcore.fitcore_solve_open(m, p(u), p(np.ascontiguousarray(pts[0])), p(np.ascontiguousarray(pts[1])),
                                              p(t2), p(cx), p(cy)) == 0
            assert np.array_equal(t, t2) and np.array_equal(c[0], cx) and np.array_equal(c[1], cy)
        x = np.concatenate([np.linspace(0, u[-1], 150), u])
        tt, cx, cy = np.ascontiguousarray(t), np.ascontiguousarray(c[0]), np.ascontiguousarray(c[1])
        for nu in (0, 1, 2):
            want = splev(x, (t, c, k), der=nu)
            ox, oy = np.zeros(x.size), np.zeros(x.size)
            fitcore.fitcore_splev(tt.size, p(tt), p(cx), p(cy), nu, p(x), x.size, p(ox), p(oy))
            assert np.array_equal(ox, want[0]) and np.array_equal(oy, want[1]), (trial, nu)
