"""CPU: the oracles against the golden vectors produced by the unmodified reference
(tools/make_golden.py).  Level 1 (Python port, same SciPy/numpy) must be bit-exact; level 2 (C, own
periodic spline) must reproduce the sweeps bit-exactly when fed the reference curvature and the whole
path to the documented noise floor ("Parity")."""
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
