"""Worker process of `TrajectoryBayesianNonlinear.optimize_COBYLA_lockstep`: scipy's COBYLA on the reference's
objective (trajectory_bayesian_nonlinear.py:207-227) with every lap time supplied by the parent process.

Started as `python -m lap_time_optimization_b200._cobyla_worker`; talks pickle over stdin/stdout:
    parent -> worker   (tau0, alpha0, maxiter)            once
    worker -> parent   ("ask", x) ... parent -> worker  tau   per objective evaluation
    worker -> parent   ("done", x_star)                   at the end
No CUDA, no torch in here."""
import pickle
import sys

import numpy as np

ALPHA_LOW, ALPHA_HIGH = 0.0, 0.99  # trajectory_bayesian_nonlinear.py:209


def main():
    from scipy.optimize import minimize

    inp, out = sys.stdin.buffer, sys.stdout.buffer
    sys.stdout = sys.stderr  # anything printed by libraries must not corrupt the protocol
    tau0, alpha0, maxiter = pickle.load(inp)

    def objective(x):
        pickle.dump(("ask", np.asarray(x, dtype=np.float64)), out)
        out.flush()
        tau = pickle.load(inp)
        return -max(0.0, tau0 - tau)  # tbn.py:211-216

    bounds = np.array([[ALPHA_LOW, ALPHA_HIGH] for _ in alpha0])
    res = minimize(objective, x0=np.asarray(alpha0, dtype=np.float64), bounds=bounds, method="COBYLA",
                   options={"maxiter": int(maxiter), "disp": False})
    pickle.dump(("done", np.asarray(res.x, dtype=np.float64)), out)
    out.flush()


if __name__ == "__main__":
    main()
