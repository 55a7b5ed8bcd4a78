#!/usr/bin/env python
"""The reference's offline trainer (GPTrainers.py:25-165) on one bundled dataset, with the three
imports swapped for gpcore and the 2000 x 2000 inversions replaced by the
device evaluator.  BASELINE.json configs[0].

    python examples/gptrainers_flow.py                 # tests/golden/field_data.npz (GPData_0.2_fieldMeas_0_T0_0)
    python examples/gptrainers_flow.py --csv GPData_x.csv --out results_dir

Prints the same RMSE / WRMSE lines the reference writes to MSE_*.txt; the published values for
this dataset are RMSE mf 5.2483, sf 5.2475, nisf 5.2474, sfTP 5.2432
(Data/TrajectoriesAndEstimates/GPResults/MSE_0.2_fieldMeas_0_T0_0.txt).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.setup_path()

import gpcore.GPy as GPy  # noqa: E402
import gpcore.emukit as emukit  # noqa: E402
from gpcore import evaluate  # noqa: E402
from gpcore.emukit.model_wrappers.gpy_model_wrappers import GPyMultiOutputWrapper  # noqa: E402
from gpcore.emukit.multi_fidelity.convert_lists_to_array import convert_xy_lists_to_arrays  # noqa: E402
from gpcore.emukit.multi_fidelity.models import GPyLinearMultiFidelityModel  # noqa: E402
from gpcore.nigp import NIGP  # noqa: E402

# Data/TrajectoriesAndEstimates/FieldData/FieldSettings0.txt
FIELD0 = dict(L=4.952356847443557, s=0.16551033487166417, w=np.array([0.19015503, 0.52624564, 1.7915839]),
              p=np.array([[2.47147946, 11.73430852, 6.50601445], [0.10847044, 4.89872175, 10.0],
                          [6.62116941, 19.86149887, 5.36276189], [4.85734684, 2.51691651, 3.0],
                          [0.85557525, 15.05903087, 9.36963751]]))


def wrbf_field(x, p, L, s, w):
    """exploreSimSettings.py:74-86 (vectorised): sum_k L exp(-(s |(x - p_k) w|)^2)."""
    d2 = (((x[:, None, :] - p[None, :, :]) * w[None, None, :]) ** 2).sum(-1) * s ** 2
    return (L * np.exp(-d2)).sum(1)[:, None]


def test_grid():
    """exploreSimSettings.py:116-119: 10 x 20 x 10 meshgrid over the workspace, ravel('F')."""
    g = np.meshgrid(np.linspace(0, 10, 10), np.linspace(0, 20, 20), np.linspace(0, 10, 10))
    return np.ascontiguousarray(np.array([gi.ravel("F") for gi in g]).T)


def run(cols, field=FIELD0, nigp_iters=10, nigp_restarts=2, verbose=True, out_dir=None, tag="GPData"):
    t_all = time.perf_counter()
    xs_true, ys = evaluate.split_fidelities(cols, estimated=False)
    xs_est, _ = evaluate.split_fidelities(cols, estimated=True)
    n_fids = 3
    Xh_train, Y_train = convert_xy_lists_to_arrays(xs_est, ys)
    kernels = [GPy.kern.RBF(3, ARD=True), GPy.kern.RBF(3, ARD=True), GPy.kern.RBF(3, ARD=True)]
    lin_mf_kernel = emukit.multi_fidelity.kernels.LinearMultiFidelityKernel(kernels)
    gpy_lin_mf_model = GPyLinearMultiFidelityModel(Xh_train, Y_train, lin_mf_kernel, n_fidelities=n_fids)
    lin_mf_model = GPyMultiOutputWrapper(gpy_lin_mf_model, n_fids, n_optimization_restarts=1)
    lin_mf_model.set_data(Xh_train, Y_train)
    lin_mf_model.gpy_model.kern.scale.fix([1, 1])
    lin_mf_model.optimize()
    emuHypVec = lin_mf_model.gpy_model.param_array.copy()

    sfXhs = np.stack([cols["xh"], cols["yh"], cols["zh"]], axis=1)
    sfXs = np.stack([cols["x"], cols["y"], cols["z"]], axis=1)
    sfys = cols["fieldVal"][:, None]
    gp = GPy.models.GPRegression(sfXhs, sfys, GPy.kern.RBF(input_dim=3, ARD=True))
    gp.set_XY(sfXhs, sfys)
    gp.optimize()
    gpTruePos = GPy.models.GPRegression(sfXs, sfys, GPy.kern.RBF(input_dim=3, ARD=True))
    gpTruePos.set_XY(sfXs, sfys)
    gpTruePos.optimize()
    nigp = NIGP(n_restarts=nigp_restarts, iters=nigp_iters, verbose=False)
    nigp.fit(sfXhs, sfys)

    testPoints = test_grid()
    fTrue = wrbf_field(testPoints, field["p"], field["L"], field["s"], field["w"])
    munisf, signisf = nigp.predict(testPoints, return_cov=1)
    musf, sigsf = gp.predict(testPoints, full_cov=1)
    musfTP, sigsfTP = gpTruePos.predict(testPoints, full_cov=1)
    t2 = np.hstack((testPoints, 2 * np.ones((testPoints.shape[0], 1))))
    mumf, sigmf = lin_mf_model.predict(t2)
    SIG = lin_mf_model.predict_covariance(t2)
    errs = {"mf": mumf - fTrue, "sf": musf - fTrue, "nisf": munisf[:, None] - fTrue, "sfTP": musfTP - fTrue}
    covs = {"mf": SIG, "sf": sigsf, "nisf": signisf, "sfTP": sigsfTP}
    rm = {k: evaluate.rmse(e) for k, e in errs.items()}
    wm = {}
    for k in errs:
        try:
            wm[k] = evaluate.weighted_mse(errs[k], covs[k])
        except np.linalg.LinAlgError:
            # emukit's element-wise 1e-10 clip makes the multi-fidelity covariance indefinite (the
            # reference inverts it by LU all the same); fall back to the un-clipped covariance
            if k != "mf":
                # a 2000 x 2000 posterior covariance that is numerically singular: the reference's LU-based
                # np.linalg.inv returns a meaningless number there (its published "WRMSE nisf" values of 1e-8 .. 1e-12
                # are of that kind); the device evaluator factors by Cholesky and says so instead
                wm[k] = float("nan")
                if verbose:
                    print("  (%s: posterior covariance is numerically not positive definite; no WRMSE)" % k)
                continue
            _, raw = lin_mf_model.gpy_model.predict(t2, full_cov=True)
            try:
                wm[k] = evaluate.weighted_mse(errs[k], raw)
            except np.linalg.LinAlgError:
                wm[k] = float("nan")
            if verbose:
                print("  (mf: clipped covariance is not positive definite; WRMSE from the un-clipped one)")
    if verbose:
        for k, v in rm.items():
            print("RMSE {}:{}".format(k, v))
        for k, v in wm.items():
            print("WRMSE {}:{}".format(k, v))
        print("hypers  mf:", np.round(emuHypVec, 4))
        print("        sf:", np.round(gp.param_array, 4), " sfTP:", np.round(gpTruePos.param_array, 4))
        print("      nisf:", np.round(nigp.get_params(), 4))
        print("wall %.1f s" % (time.perf_counter() - t_all))
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        np.savetxt(os.path.join(out_dir, tag + "_emuGP.txt"), emuHypVec[None, :], delimiter=",")
        np.savetxt(os.path.join(out_dir, tag + "_sfGP.txt"), gp.param_array, delimiter=",")
        np.savetxt(os.path.join(out_dir, tag + "_sfGPTP.txt"), gpTruePos.param_array, delimiter=",")
        np.savetxt(os.path.join(out_dir, tag + "_nisfGP.txt"), nigp.get_params(), delimiter=",")
        evaluate.write_gpres_csv(os.path.join(out_dir, tag.replace("GPData", "GPRes") + ".csv"), testPoints, fTrue, musf,
                                 np.diag(sigsf)[:, None], mumf, sigmf)
        evaluate.write_mse_txt(os.path.join(out_dir, tag.replace("GPData", "MSE") + ".txt"), rm, wm)
    return rm, wm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--csv", default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--nigp-iters", type=int, default=10)
    a = ap.parse_args()
    if a.csv:
        cols = evaluate.read_gpdata_csv(a.csv)
        tag = os.path.basename(a.csv).replace(".csv", "")
    else:
        d = np.load(os.path.join(ROOT, "tests", "golden", "field_data.npz"))
        keep = d["t"] < 3600
        cols = {"t": d["t"][keep], "x": d["X"][keep, 0], "y": d["X"][keep, 1], "z": d["X"][keep, 2],
                "xh": d["Xh"][keep, 0], "yh": d["Xh"][keep, 1], "zh": d["Xh"][keep, 2],
                "fieldVal": d["y"][keep], "fidLev": d["fidLev"][keep]}
        tag = "GPData_0.2_fieldMeas_0_T0_0"
    np.random.seed(0)
    run(cols, nigp_iters=a.nigp_iters, out_dir=a.out, tag=tag)


if __name__ == "__main__":
    main()
