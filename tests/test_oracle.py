"""CPU: the oracles against the golden vectors produced by the unmodified reference
(tools/make_golden.py).  Level 1 (Python port, same SciPy/numpy) must be bit-exact; level 2 (C, own
periodic spline) must reproduce the sweeps bit-exactly when fed the reference curvature and the whole
path to the documented noise floor ("Parity")."""
import os

import numpy as np
import pytest

from conftest import case_setup, golden_cases, rel_err
from oracle import c_oracle
from oracle.reference_port import OracleEvaluator, OracleTrack, load_vehicle, top_k

PROFILE_KEYS = ("k", "v_local", "v_acclim", "v_declim", "v")


def _port(name, ns):
    tj, width, vj, mode = case_setup(name)
    return OracleEvaluator(OracleTrack(tj, width), load_vehicle(vj), mode, ns)


def _c(name, ns, **kw):
    tj, width, vj, mode = case_setup(name)
    return c_oracle.COracle(OracleTrack(tj, width), load_vehicle(vj), mode, ns, **kw)


@pytest.mark.parametrize("name", golden_cases())
def test_port_bit_exact(name, golden):
    g = golden(name)
    ns = int(g["ns"])
    ev = _port(name, ns)
    n = min(len(g["alphas"]), 12 if ns < 2000 else 2)
    for i in range(n):
        assert ev.lap_time(g["alphas"][i]) == g["laps"][i]
    for i in range(min(int(g["n_profiles"]), 2)):
        pr = ev.profile(g["alphas"][i])
        for key in PROFILE_KEYS + ("s", "controls"):
            assert np.array_equal(pr[key], g["prof_" + key][i]), key
        assert pr["length"] == g["prof_length"][i]


def test_known_answers(golden):
    """SURVEY.md section 8(c) known-answer values."""
    g = golden("buckmore_tbr18_bayes_known")
    assert g["laps"][0] == 45.16138534803076
    assert golden("buckmore_tbr18_full")["laps"][0] == 47.03786396842785  # centre line, TBR18
    assert golden("buckmore_mx5_full")["laps"][0] == 59.89998713338799  # centre line, MX5


@pytest.mark.parametrize("name", golden_cases())
def test_c_sweeps_bit_exact_given_reference_curvature(name, golden):
    g = golden(name)
    co = _c(name, int(g["ns"]), use_pow=True)
    for i in range(int(g["n_profiles"])):
        sw = co.sweeps(g["prof_k"][i], g["prof_length"][i])
        for key in ("v_local", "v_acclim", "v_declim", "v"):
            assert np.array_equal(sw[key], g["prof_" + key][i]), key
        assert sw["lap"] == g["prof_lap"][i]
    # x*x instead of libm pow(x, 2): what the CUDA kernels do.  1-ulp differences in ~0.08 % of squares.
    co2 = _c(name, int(g["ns"]))
    for i in range(int(g["n_profiles"])):
        sw = co2.sweeps(g["prof_k"][i], g["prof_length"][i])
        assert abs(sw["lap"] - g["prof_lap"][i]) <= 1e-12 * g["prof_lap"][i]


@pytest.mark.parametrize("name", golden_cases())
def test_c_full_path(name, golden):
    g = golden(name)
    co = _c(name, int(g["ns"]))
    laps = co.lap_times(g["alphas"])
    rel = rel_err(laps, g["laps"])
    for i in range(int(g["n_profiles"])):
        pr = co.profile(g["alphas"][i])
        assert pr["length"] == g["prof_length"][i]
        # curvature of the closed-form periodic spline vs FITPACK's
        assert np.max(np.abs(pr["k"] - g["prof_k"][i])) <= 1e-12 * np.max(g["prof_k"][i])
    if "mx5" in name:
        assert rel.max() <= 1e-13  # no cancellation in the MX5 traction
    elif "full" in name:
        # random alphas on every cone give zig-zag lines (laps up to 200 s) that ride the friction
        # limit for most of the lap: the TBR18 sqrt(f^2 - f_lat^2) noise floor is ~1e-8 there
        assert np.median(rel) <= 1e-9 and rel.max() <= 1e-7
    else:
        assert np.median(rel) <= 1e-10 and rel.max() <= 1e-9


def test_c_topk_matches_port_population(golden):
    g = golden("buckmore_tbr18_bayes")
    co = _c("buckmore_tbr18_bayes", int(g["ns"]))
    laps = co.lap_times(g["alphas"])
    i_ref, _ = top_k(list(g["laps"]), 10)
    i_c, _ = top_k(list(laps), 10)
    assert np.array_equal(i_ref, i_c)


def test_pairwise_sum_is_numpy_sum():
    rng = np.random.default_rng(3)
    for n in (1, 7, 8, 100, 128, 129, 846, 2500, 10000):
        a = rng.uniform(0.01, 0.2, n)
        assert c_oracle.pairwise_sum(a) == np.sum(a)


def test_topk_stable():
    laps = [3.0, 1.0, 2.0, 1.0, 5.0, 2.0]
    idx, best = top_k(laps, 4)
    assert list(idx) == [1, 3, 2, 5] and list(best) == [1.0, 1.0, 2.0, 2.0]


# ---- FITPACK-faithful spline (oracle/fitpack_port.c) ------------------------------------------------
def _closed_polygon(rng, m):
    th = np.sort(rng.uniform(0, 2 * np.pi, m - 1))
    r = rng.uniform(50, 120, m - 1)
    pts = np.array([r * np.cos(th), r * np.sin(th)])
    pts = np.concatenate([pts, pts[:, :1]], axis=1)
    u = np.append(0, np.cumsum(np.linalg.norm(np.diff(pts, axis=1), axis=0)))
    return u, pts


def test_fitpack_port_equals_live_scipy():
    """fpclos (s = 0, per = 1) and splder restated in C against the installed SciPy, bit for bit:
    the call the reference makes at path.py:25 and path.py:51-54."""
    from scipy.interpolate import splev, splprep

    rng = np.random.default_rng(11)
    for trial in range(60):
        m = int(rng.integers(6, 160))
        u, pts = _closed_polygon(rng, m)
        (t, c, k), _ = splprep(pts.copy(), u=u, k=3, s=0, per=1)
        t2, c2 = c_oracle.fitpack_spline(u, pts)
        assert np.array_equal(t, t2)
        assert np.array_equal(c[0], c2[0]) and np.array_equal(c[1], c2[1])
        x = np.linspace(0, u[-1], 401)[:-1]
        for nu in (0, 1, 2):
            ys = splev(x, (t, c, k), der=nu)
            for d in range(2):
                assert np.array_equal(ys[d], c_oracle.fitpack_splder(t2, c2[d], nu, x)), (trial, nu, d)


def test_fitpack_port_open_spline_equals_live_scipy():
    """fppara (s = 0, per = 0: the not-a-knot interpolating spline of open paths) restated in C against SciPy."""
    from scipy.interpolate import splev, splprep

    rng = np.random.default_rng(12)
    for trial in range(40):
        m = int(rng.integers(4, 140))
        th = np.sort(rng.uniform(0, 1.7 * np.pi, m))
        r = rng.uniform(50, 120, m)
        pts = np.array([r * np.cos(th), r * np.sin(th)])
        u = np.append(0, np.cumsum(np.linalg.norm(np.diff(pts, axis=1), axis=0)))
        (t, c, k), _ = splprep(pts.copy(), u=u, k=3, s=0, per=0)
        t2, c2 = c_oracle.fitpack_spline_open(u, pts)
        assert np.array_equal(t, t2) and np.array_equal(c[0], c2[0]) and np.array_equal(c[1], c2[1])
        x = np.linspace(0, u[-1], 201)
        for nu in (0, 1, 2):
            ys = splev(x, (t, c, k), der=nu)
            for d in range(2):
                assert np.array_equal(ys[d], c_oracle.fitpack_splder(t2, c2[d], nu, x)), (trial, nu, d)


@pytest.mark.parametrize("name", golden_cases())
def test_fitpack_port_equals_reference_spline(name, golden):
    """Same check against what the unmodified reference held in Path.spline (tools/make_golden.py)."""
    g = golden(name)
    for i in range(int(g["n_profiles"])):
        t2, c2 = c_oracle.fitpack_spline(g["prof_dists"][i], g["prof_controls"][i])
        assert np.array_equal(t2, g["prof_tck_t"][i])
        assert np.array_equal(c2[0], g["prof_tck_cx"][i]) and np.array_equal(c2[1], g["prof_tck_cy"][i])
        x = g["prof_s"][i][:-1]
        for key, d, nu in (("dx", 0, 1), ("dy", 1, 1), ("ddx", 0, 2), ("ddy", 1, 2)):
            assert np.array_equal(c_oracle.fitpack_splder(t2, c2[d], nu, x), g["prof_" + key][i]), key


def test_pow15_is_correctly_rounded():
    """x**1.5 by two double-double steps against exact integer arithmetic."""
    from fractions import Fraction
    from math import isqrt

    rng = np.random.default_rng(5)
    for x in np.concatenate([rng.uniform(0.3, 3.0, 1500), rng.uniform(1e-3, 1e3, 500)]):
        x = float(x)
        got = c_oracle.pow15(x)
        f = Fraction(x) ** 3
        root = Fraction(isqrt((f.numerator << 600) // f.denominator), 1 << 300)  # sqrt(x^3), 300 bits
        cands = [got, float(np.nextafter(got, np.inf)), float(np.nextafter(got, -np.inf))]
        assert min(cands, key=lambda v: abs(Fraction(v) - root)) == got


@pytest.mark.parametrize("name", golden_cases())
def test_c_fitpack_mode_reproduces_reference(name, golden):
    """Whole path with FITPACK's arithmetic.  Against the reference on numpy's baseline dispatch (libm
    pow) the laps are bit-equal on nearly every candidate; against the reference on an AVX512 host (numpy's
    SVML pow: one ulp off in ~5 % of the curvature samples) they carry that host's own spread."""
    g = golden(name)
    co = _c(name, int(g["ns"]), spline="fitpack", use_pow=True)
    laps = co.lap_times(g["alphas"])
    rel_base = rel_err(laps, g["laps_base"])
    rel_host = rel_err(laps, g["laps"])
    spread = rel_err(g["laps_base"], g["laps"]).max()  # the reference against itself
    assert (rel_base == 0).mean() >= 0.8, (rel_base == 0).mean()
    assert rel_base.max() <= 1e-10
    assert rel_host.max() <= max(1e-9, 2 * spread)
    for i in range(int(g["n_profiles"])):
        pr = co.profile(g["alphas"][i])
        assert pr["length"] == g["prof_length"][i]
        kb = g["prof_k_base"][i]
        assert (pr["k"] != kb).mean() <= 0.005  # libm pow is not always correctly rounded either
        ulp = np.spacing(np.maximum(pr["k"], kb))
        # one ulp of the power where they differ = up to two ulp of the quotient across a binade
        assert np.all(np.abs(pr["k"] - kb) <= 2 * ulp)
        assert np.all(np.abs(pr["k"] - g["prof_k"][i]) <= 3 * ulp)


def test_reference_arithmetic_depends_on_numpy_dispatch():
    """The port under numpy's baseline dispatch (libm pow) against the port under this host's dispatch: identical on
    hosts without AVX512, within ~1e-8 otherwise (numpy's SVML `x ** 1.5` differs from libm's by 1 ulp in ~5 % of the
    arguments; the friction-circle cancellation of TBR18 amplifies it).  This is the noise floor of the reference
    against itself that the 1e-9 parity budget has to be read against."""
    from oracle.reference_port import lap_times_baseline_dispatch, lap_times_pool

    tj, width, vj, mode = case_setup("buckmore_tbr18_bayes")
    a = np.random.default_rng(3).uniform(0.0, 0.99, (96, 43))
    host = lap_times_pool(tj, width, vj, a, mode, processes=2)
    base = lap_times_baseline_dispatch(tj, width, vj, a, mode)
    rel = np.abs(host - base) / base
    assert rel.max() < 2e-8


@pytest.mark.parametrize("veh", ["tbr18", "mx5"])
@pytest.mark.parametrize("mode", ["bayes", "full"])
def test_plateau_circle_port_equals_reference(tmp_path, veh, mode):
    """Curvature plateaus (circular corridor, symmetric candidates; tools/make_golden_plateau.py ran the unmodified
    reference): the port reproduces the reference's lap times bit for bit there too, the C oracle to tolerance."""
    import lap_time_optimization_b200 as ltk
    from conftest import GOLDEN_DIR, circle_track_json
    from oracle import c_oracle
    from oracle.reference_port import OracleEvaluator, OracleTrack, load_vehicle

    g = dict(np.load(os.path.join(GOLDEN_DIR, "plateau_circle.npz")))
    tj = circle_track_json(tmp_path, g)
    vj = ltk.data_path("vehicles", "MX5.json" if veh == "mx5" else "tbr18.json")
    a, laps = g[f"{veh}_{mode}_alphas"], g[f"{veh}_{mode}_laps"]
    port = OracleEvaluator(OracleTrack(tj, float(g["width"])), load_vehicle(vj), mode)
    got = np.array([port.lap_time(x) for x in a[:24]])
    assert np.array_equal(got, laps[:24])
    for spline, tol in (("fitpack", 5e-10), ("tridiagonal", 5e-9)):  # (golden = the reference on an AVX512 host: its own pow noise)
        co = c_oracle.COracle(OracleTrack(tj, float(g["width"])), load_vehicle(vj), mode, None, spline=spline)
        assert rel_err(co.lap_times(a), laps).max() <= tol, spline

