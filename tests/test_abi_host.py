"""CPU: the C-ABI library loads and exports every symbol include/gpcore.h declares, the product
fails loudly without a CUDA device (no CPU fallback), and the host-side mirror logic
(parameter views, list conversion, fidelity labels, sharding arithmetic) behaves like the
reference's objects."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gpcore.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    from gpcore import _lib
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(built_lib, n), "libgpcore.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert built_lib.gpc_version() >= 100


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "adaptive-exploration-under-localization-uncertainty-using-multi-fidelity-"
                             "gaussian-processes_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".inc")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "/root/reference" not in txt, f


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(built_lib):
    import gpcore
    with pytest.raises(gpcore.GpcoreError, match="no CUDA device|CUDA"):
        gpcore.GPCore(0, 1, 0)
    from gpcore.GPy.kern import RBF
    from gpcore.GPy.models import GPRegression
    m = GPRegression(np.zeros((4, 3)), np.zeros((4, 1)), RBF(3, ARD=True))   # host state only
    with pytest.raises(gpcore.GpcoreError):
        m.predict(np.zeros((2, 3)))
    # GPy's keyword signature as HowManyPoints.py:91-92 calls it; anything but the zero mean is refused loudly
    GPRegression(np.zeros((4, 3)), np.zeros((4, 1)), RBF(input_dim=3, variance=1, lengthscale=1, ARD=True), mean_function=None)
    with pytest.raises(NotImplementedError):
        GPRegression(np.zeros((4, 3)), np.zeros((4, 1)), RBF(3), mean_function=object())


def test_param_array_layout_and_views():
    from gpcore.GPy.kern import RBF, Matern32
    from gpcore.GPy.likelihoods import Gaussian
    from gpcore.GPy.models import GPRegression
    from gpcore.emukit.multi_fidelity.kernels import LinearMultiFidelityKernel
    from gpcore.emukit.multi_fidelity.models import GPyLinearMultiFidelityModel
    from gpcore.emukit.model_wrappers.gpy_model_wrappers import GPyMultiOutputWrapper
    gp = GPRegression(np.zeros((5, 3)), np.zeros((5, 1)), RBF(input_dim=3, variance=2.0, lengthscale=[1, 2, 3], ARD=True))
    assert np.allclose(gp.param_array, [2, 1, 2, 3, 1])              # ...SFGP.py:620 order
    gp.Gaussian_noise.variance = 0.25                                   # informationGainTest.py:24
    assert gp.param_array[-1] == 0.25 and gp.Gaussian_noise.variance[0] == 0.25
    gp.param_array[:-1] = [9, 8, 7, 6]                                  # ...MFGP.py:416
    assert gp.kern.variance[0] == 9 and list(gp.kern.lengthscale) == [8, 7, 6]
    gp.param_array[gp.param_array > 8.5] = 1                            # ...SFGP.py:382
    assert gp.kern.variance[0] == 1
    assert gp.parameter_names() == ["rbf.variance", "rbf.lengthscale", "Gaussian_noise.variance"]
    c = gp.copy()
    c.param_array[0] = 5
    assert gp.param_array[0] == 1 and c.kern.variance[0] == 5

    X4 = np.hstack([np.zeros((6, 3)), np.array([[0, 0, 1, 1, 2, 2]]).T])
    for kern_cls, names0 in ((RBF, "multifidelity.rbf.variance"), (Matern32, "multifidelity.Mat32.variance")):
        k = LinearMultiFidelityKernel([kern_cls(3, ARD=True) for _ in range(3)])
        m = GPyLinearMultiFidelityModel(X4, np.zeros((6, 1)), k, n_fidelities=3)
        assert m.param_array.size == 17 and m.parameter_names()[0] == names0     # MixedNoise: 3 variances
        w = GPyMultiOutputWrapper(m, 3, n_optimization_restarts=1)
        w.gpy_model.kern.scale.fix([1, 1])                                    # GPTrainers.py:67
        m.param_array[:] = np.arange(1, 18)
        assert m.kern.scale[0] == 13 and m.kern.kernels[1].lengthscale[2] == 8
        assert np.sum(m.param_array[[0, 4, 8, -1]]) == 1 + 5 + 9 + 17        # exploreSimSettings.py:16
        free, _ = m._free_mask()
        assert not free[12] and not free[13] and free[:12].all()
    k = LinearMultiFidelityKernel([RBF(3, ARD=True) for _ in range(3)])
    m15 = GPyLinearMultiFidelityModel(X4, np.zeros((6, 1)), k, likelihood=Gaussian(), n_fidelities=3)
    assert m15.param_array.size == 15                                       # ...MFGP.py:658-660
    k.rbf.lengthscale.constrain_bounded(0.0001, 100)                        # ...MFGP.py:664
    assert k.rbf_1 is k.kernels[1] and k.rbf_2 is k.kernels[2]
    with pytest.raises(ValueError):
        GPyLinearMultiFidelityModel(np.hstack([np.zeros((2, 3)), [[0], [3]]]), np.zeros((2, 1)), k, n_fidelities=3)


def test_convert_lists_and_fidelity_labels():
    from gpcore.emukit.multi_fidelity.convert_lists_to_array import convert_xy_lists_to_arrays
    from gpcore.infogain import label_fidelity
    xs = [np.ones((2, 3)), 2 * np.ones((3, 3)), np.zeros((0, 3))]
    ys = [np.ones((2, 1)), np.ones((3, 1)), np.zeros((0, 1))]
    X, Y = convert_xy_lists_to_arrays(xs, ys)
    assert X.shape == (5, 4) and list(X[:, 3]) == [0, 0, 1, 1, 1] and Y.shape == (5, 1)
    with pytest.raises(ValueError):
        convert_xy_lists_to_arrays(xs[:2], ys)
    fl = [0.25, 2.25, 6.25]
    v = np.array([0.1, 0.25, 1.0, 2.25, 5.0, 9.0])
    # strict inequalities: exact ties fall to index 0 (GraceRIGV3.py:529-533)
    assert list(label_fidelity(v, fl)) == [2, 0, 1, 0, 0, 0]


def test_shard_arithmetic():
    from gpcore.sharding import shard_candidates, shard_range
    for total in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    offs = np.array([0, 3, 3, 10, 12, 20])
    lo, hi, loc, r0, r1 = shard_candidates(offs, 1, 2)
    assert (lo, hi) == (3, 5) and list(loc) == [0, 2, 10] and (r0, r1) == (10, 20)


def test_to_x4_and_ragged():
    from gpcore import GPCore, to_x4
    a = to_x4(np.array([[1.0], [2.0]]))
    assert a.shape == (2, 4) and a[1, 0] == 2 and a[:, 1:].sum() == 0
    with pytest.raises(ValueError):
        to_x4(np.zeros((2, 4)))
    rows, offs = GPCore._ragged([np.ones((2, 4)), np.zeros((0, 4)), 2 * np.ones((3, 4))])
    assert list(offs) == [0, 2, 2, 5] and rows.shape == (5, 4) and rows[2, 0] == 2


def test_trainer_io_formats(tmp_path):
    """GPData CSV reader / fidelity split / result writers (GPTrainers.py:29-61,138-165)."""
    from gpcore import evaluate
    d = np.load(os.path.join(ROOT, "tests", "golden", "field_data.npz"))
    cols = ["t", "x", "y", "z", "xh", "yh", "zh", "fieldVal", "fidLev"]
    tab = np.column_stack([d["t"], d["X"], d["Xh"], d["y"], d["fidLev"]])
    p = tmp_path / "GPData_test.csv"
    np.savetxt(p, tab, delimiter=",", header=",".join(cols), comments="")
    got = evaluate.read_gpdata_csv(str(p))
    assert list(got) == cols and np.allclose(got["xh"], d["Xh"][:, 0]) and len(got["t"]) == np.sum(d["t"] < 3600)
    xs, ys = evaluate.split_fidelities(got)
    assert [len(x) for x in xs] == [int(np.sum(got["fidLev"] == lv)) for lv in (3, 2, 1)]
    assert xs[0].shape[1] == 3 and ys[2].shape[1] == 1
    from gpcore.emukit.multi_fidelity.convert_lists_to_array import convert_xy_lists_to_arrays
    X4, Y = convert_xy_lists_to_arrays(xs, ys)
    assert X4.shape == (len(got["t"]), 4) and set(np.unique(X4[:, 3])) <= {0.0, 1.0, 2.0}
    m = tmp_path / "MSE_test.txt"
    evaluate.write_mse_txt(str(m), {"mf": 5.2483, "sf": 5.2475}, {"mf": np.array([[0.1]]), "sf": 0.2})
    back = evaluate.read_mse_txt(str(m))
    assert back["RMSE mf"] == 5.2483 and back["WRMSE mf"] == 0.1 and back["WRMSE sf"] == 0.2
    tp = d["test_sub"]
    evaluate.write_gpres_csv(str(tmp_path / "GPRes.csv"), tp, tp[:, :1], tp[:, :1], tp[:, :1], tp[:, :1], tp[:, :1])
    assert np.loadtxt(tmp_path / "GPRes.csv", delimiter=",", skiprows=1).shape == (len(tp), 8)


def test_parameter_transforms_softplus_and_logistic():
    """optimize() works in GPy's raw spaces: Logexp (softplus) for positives, Logistic for ``constrain_bounded`` --
    smooth, strictly inside the bounds, derivative consistent with the map (a clamp made the objective
    flat past a bound while the gradient stayed non-zero)."""
    from gpcore.gp_models import _transforms
    lo = np.array([0.0, 1e-4, 0.0, 0.5])
    hi = np.array([np.inf, 100.0, np.inf, 2.0])
    from_raw, dth, to_raw = _transforms(lo, hi)
    th0 = np.array([2.5, 3.0, 1e-3, 1.25])
    x0 = to_raw(th0)
    assert np.allclose(from_raw(x0), th0, rtol=1e-12, atol=0)
    for x in (x0, x0 + 3.0, x0 - 5.0, np.array([40.0, 50.0, -30.0, -50.0])):
        th = from_raw(x)
        assert np.all(th[[1, 3]] >= lo[[1, 3]]) and np.all(th[[1, 3]] <= hi[[1, 3]]) and np.all(th > 0)
        e = 1e-6
        fd = (from_raw(x + e) - from_raw(x - e)) / (2 * e)
        assert np.allclose(dth(th), fd, rtol=1e-5, atol=1e-12)
    # a start on or outside a bound is moved strictly inside instead of producing an infinite raw value
    assert np.all(np.isfinite(to_raw(np.array([1.0, 100.0, 1.0, 0.5]))))
    assert np.all(np.isfinite(to_raw(np.array([1.0, 500.0, 1.0, 0.1]))))


def test_multi_fidelity_rows_are_validated_on_the_host():
    from gpcore.gp_models import _x4_mf
    ok = np.array([[0.0, 1.0, 2.0, 1.0], [3.0, 4.0, 5.0, 0.0]])
    assert _x4_mf(ok, 2) is ok
    for bad in (2.0, -1.0, 0.5, np.nan):
        X = ok.copy()
        X[1, 3] = bad
        with pytest.raises(ValueError, match="fidelity"):
            _x4_mf(X, 2)
    assert _x4_mf(np.array([[1.0, 2.0, 1.0]]), 3).shape == (1, 4)       # (x, y, fid) -> (x, y, 0, fid)
