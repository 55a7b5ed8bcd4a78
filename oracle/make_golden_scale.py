"""Golden vectors at the BASELINE sizes, produced by the reference's OWN code.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs ``/root/reference``):

    python oracle/make_golden_scale.py

* ``nigp_8192.npz`` -- BASELINE configs[2] training set exactly as ``bench.py`` builds it
  (``make_train(8192, 3)``, fidelity column zeroed, hypers ``bench.NIGP_HYP``): the reference module
  ``NIGP.py`` -- imported VERBATIM through ``oracle/gpy_shim`` -- computes the posterior-mean
  gradients of the training inputs (``compute_post_mean_and_gradients``, ``NIGP.py:29-65``), the
  per-point noise ``sum_d g_d^2 sigma_x_d^2`` (``NIGP.py:251-252``) and ``NIGP.predict``
  (``NIGP.py:269-333``) on 2000 test points (500 of them 1e-3 away from training inputs, where the
  variance cancels to ~noise), with and without ``Xs_input_noise``.  The training inputs are NOT
  stored (they are regenerated from the seed; a checksum is); the test points and outputs are.
  ``spread_*``: the same quantities from the restatement with direct-difference distances
  (``gp_oracle.nigp_predict(gram=False)``) against the reference's Gram-trick arithmetic -- the
  CPU-vs-CPU formulation spread that bounds how closely ANY implementation can follow the reference
  at this condition number.
* ``scale_oracle.npz`` -- see ``scale_oracle()``.
* ``nigp_fit.npz`` -- see ``nigp_fit_golden()``.
* ``gp_datasets.npz`` -- see ``datasets_golden()``.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def reference_nigp():
    sys.path.insert(0, os.path.join(HERE, "gpy_shim"))
    sys.path.insert(0, REF)
    import NIGP as ref
    assert os.path.realpath(ref.__file__).startswith(REF), ref.__file__
    return ref


def nigp_8192(N=8192, M=2000):
    import bench
    from oracle import gp_oracle as go
    ref = reference_nigp()
    hyp = bench.NIGP_HYP
    X4, y = bench.make_train(N, 3)
    X = np.ascontiguousarray(X4[:, :3])
    rng = np.random.default_rng(8192)
    Xs = rng.uniform([0, 0, 0], [10, 20, 10], (M, 3))
    Xs[:500] = X[rng.choice(N, 500, replace=False)] + 1e-3
    t0 = time.time()
    fm, grads = ref.compute_post_mean_and_gradients(X, y, hyp["ls"], hyp["sigma_f"], hyp["sigma_y"])
    noise_diag = np.sum(grads ** 2 * hyp["sigma_x"][None, :] ** 2, axis=1)
    m = ref.NIGP(verbose=False)
    m.lengthscales_, m.sigma_f_, m.sigma_y_, m.sigma_x_ = hyp["ls"], hyp["sigma_f"], hyp["sigma_y"], hyp["sigma_x"]
    m.X_train_, m.y_train_, m.noise_diag_train_ = X, y, noise_diag
    mean, var = m.predict(Xs)
    _, var_in = m.predict(Xs, Xs_input_noise=hyp["sigma_x"])
    print("reference NIGP.py: %.1f s" % (time.time() - t0))
    # CPU-vs-CPU spread: direct differences instead of the Gram trick, same LAPACK
    mean_d, var_d = go.nigp_predict(X, y, hyp["ls"], hyp["sigma_f"], hyp["sigma_y"], noise_diag, Xs, gram=False)
    sf = hyp["sigma_f"]
    spread_mean = float(np.max(np.abs(mean_d - mean)) / np.max(np.abs(mean)))
    spread_var = float(np.max(np.abs(var_d - var)) / sf)
    spread_var_elem = float(np.max(np.abs(var_d - var) / np.abs(var)))
    print("spread gram vs direct: mean %.2e var %.2e (normwise) %.2e (element-wise)" % (spread_mean, spread_var, spread_var_elem))
    sha = hashlib.sha256(X.tobytes() + y.tobytes()).hexdigest()
    np.savez_compressed(os.path.join(OUT, "nigp_8192.npz"), N=N, train_sha256=sha, Xs=Xs, noise_diag=noise_diag,
                        f_mean_train_head=fm[:256], grads_head=grads[:256], mean=mean, var=var, var_in=var_in,
                        mean_direct=mean_d, var_direct=var_d,
                        spread_mean=spread_mean, spread_var=spread_var, spread_var_elem=spread_var_elem)


def nigp_fit_golden():
    """``nigp_fit.npz``: the reference's own ``NIGP.fit`` (``NIGP.py:191-260``) on the seed-0 data of nigp_demo.npz with
    the ``__main__`` protocol (n_restarts = 2, iters = 10; the restart perturbations come from NumPy's global
    generator, which at that point has produced the 40 + 40 normals of the data set): fitted hypers, per-point noise and
    the objective at the fitted point.  L-BFGS-B differentiates the objective NUMERICALLY (step 1e-8), so the fitted
    hypers are reproducible only as far as rounding noise in the objective allows: ``spread_params`` is the relative
    distance between the reference's result and the same loop over the restated objective with Gram-trick and with
    direct-difference distances (objective perturbations of ~1e-13) -- the floor of any fit-parity tolerance."""
    from oracle import gp_oracle as go
    ref = reference_nigp()
    g = np.load(os.path.join(OUT, "nigp_demo.npz"))
    X, y = g["X"], g["y"]

    def seeded():
        np.random.seed(0)
        np.random.randn(40, 1)
        np.random.randn(40)

    seeded()
    m = ref.NIGP(n_restarts=2, iters=10, verbose=False)
    m.fit(X, y)
    params = m.get_params()
    assert np.allclose(params, g["params"], rtol=1e-12), "the committed demo golden was fitted by this protocol"
    log_hyp = np.log(np.concatenate([m.lengthscales_, [m.sigma_f_, m.sigma_y_], m.sigma_x_]))
    nlml_fit = go.nigp_nlml(log_hyp, X, y, np.zeros_like(X), m.noise_diag_train_)
    out = dict(X=X, y=y, n_restarts=2, iters=10, maxiter_opt=200, params=params, log_hyp=log_hyp,
               noise_diag=m.noise_diag_train_, nlml_at_fit=nlml_fit)
    spread = 0.0
    for gram in (True, False):
        seeded()
        p, _, _, bv = go.nigp_fit(X, y, 2, 10, 200, gram=gram)
        out["params_restated_gram%d" % gram] = p
        out["nlml_restated_gram%d" % gram] = bv
        spread = max(spread, float(np.max(np.abs(p - params) / np.abs(params))))
    out["spread_params"] = spread
    print("nigp_fit: params", params, "nlml", nlml_fit, "CPU-vs-CPU spread %.2e" % spread)
    np.savez_compressed(os.path.join(OUT, "nigp_fit.npz"), **out)


def datasets_golden():
    """``gp_datasets.npz``: twelve of the reference's bundled 709-row data sets (field 0, trajectories T0..T6, mixed
    localisation-noise suffixes) with the RMSE / WRMSE numbers the
    reference PUBLISHED for each (Data/TrajectoriesAndEstimates/GPResults/MSE_*.txt), and the two field definitions."""
    import ast
    import re
    base = os.path.join(REF, "Data", "TrajectoriesAndEstimates")
    # field 0 only: for the field-5 family the published RMSEs of ALL four models (the GPy-free NIGP included) are off
    # by one common factor per data set from what the bundled FieldSettings5.txt gives -- those files were produced
    # with another truth field and pin nothing
    picks = [(0, t, sfx) for t, sfx in ((0, "0"), (0, "0.1"), (1, "0.1"), (1, "0.2"), (2, "0.2"), (2, "0"), (3, "0.1"), (3, "0"),
                                        (4, "0"), (4, "0.2"), (5, "0.2"), (6, "0"))]
    out = {"names": []}
    for f, t, sfx in picks:
        name = "0.2_fieldMeas_%d_T%d_%s" % (f, t, sfx)
        path = os.path.join(base, "GPDataSets", "GPData_%s.csv" % name)
        with open(path) as fh:
            hdr = fh.readline().strip().split(",")
            d = np.loadtxt(fh, delimiter=",")
        cols = ["t", "x", "y", "z", "xh", "yh", "zh", "fieldVal", "fidLev"]
        out["data_" + name] = np.stack([d[:, hdr.index(c)] for c in cols], axis=1)
        pub = {}
        for line in open(os.path.join(base, "GPResults", "MSE_%s.txt" % name)):
            m = re.match(r"(W?RMSE) (\w+):\[*([-+0-9.eE]+)", line.strip())
            if m:
                pub[m.group(1) + "_" + m.group(2)] = float(m.group(3))
        out["pub_" + name] = np.array([pub[k] for k in ("RMSE_mf", "RMSE_sf", "RMSE_nisf", "RMSE_sfTP", "WRMSE_mf", "WRMSE_sf",
                                                        "WRMSE_nisf", "WRMSE_sfTP")])
        out["names"].append(name)
    out["names"] = np.array(out["names"])
    out["columns"] = np.array(["t", "x", "y", "z", "xh", "yh", "zh", "fieldVal", "fidLev"])
    for f in (0,):
        txt = open(os.path.join(base, "FieldData", "FieldSettings%d.txt" % f)).read()
        m = re.search(r"L,s,w: \(([-0-9.eE]+), ([-0-9.eE]+), array\(\[([^\]]+)\]\)\)", txt)
        src = re.search(r"sources:\s*\[\[(.*?)\]\]", txt, re.S).group(1)
        rows = [[float(v) for v in r.replace("[", "").replace("]", "").split()] for r in src.split("\n")]
        out["field%d_Lsw" % f] = np.array([float(m.group(1)), float(m.group(2))] + [float(v) for v in m.group(3).split(",")])
        out["field%d_p" % f] = np.array(rows)
    np.savez_compressed(os.path.join(OUT, "gp_datasets.npz"), **out)
    print("gp_datasets:", list(out["names"]), out["field0_p"].shape, os.path.getsize(os.path.join(OUT, "gp_datasets.npz")))


def scale_oracle():
    """``scale_oracle.npz``: the oracle's outputs on the seeded BASELINE-size cases of tests/scale_cases.py
    (configs[1] N = 2048 posterior, configs[3] N = 4096 information gain by the LITERAL refit loops, N = 16384 single
    fidelity), each keyed by a checksum of its inputs."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scale_cases as sc
    out = {}
    for fn in (sc.configs1_oracle, sc.configs3_oracle, sc.sf16384_oracle):
        t0 = time.time()
        out.update(fn())
        print(fn.__name__, "%.1f s" % (time.time() - t0))
    np.savez_compressed(os.path.join(OUT, "scale_oracle.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1:] or ["nigp", "scale", "fit"]
    if "fit" in what:
        nigp_fit_golden()
    if "datasets" in what:
        datasets_golden()
    if "nigp" in what:
        nigp_8192()
    if "scale" in what:
        scale_oracle()
