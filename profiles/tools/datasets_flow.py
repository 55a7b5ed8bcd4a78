"""BASELINE configs[0] over twelve bundled data sets: the reference's GPTrainers.py flow (examples/gptrainers_flow.py) on each
data set of tests/golden/gp_datasets.npz, next to the RMSE / WRMSE numbers the reference published for it.
    python profiles/tools/datasets_flow.py > gpurun_out/datasets_flow.jsonl"""
import importlib.util, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spec = importlib.util.spec_from_file_location("gptrainers_flow", os.path.join(ROOT, "examples", "gptrainers_flow.py"))
flow = importlib.util.module_from_spec(spec); spec.loader.exec_module(flow)
g = np.load(os.path.join(ROOT, "tests", "golden", "gp_datasets.npz"))
cols_names = [str(c) for c in g["columns"]]
for name in [str(n) for n in g["names"]]:
    d = g["data_" + name]
    d = d[d[:, 0] < 3600]
    cols = {c: d[:, i] for i, c in enumerate(cols_names)}
    fid = 0
    Lsw = g["field%d_Lsw" % fid]
    field = dict(L=Lsw[0], s=Lsw[1], w=Lsw[2:], p=g["field%d_p" % fid])
    np.random.seed(0)
    t0 = time.time()
    rm, wm = flow.run(cols, field=field, verbose=False)
    pub = g["pub_" + name]
    keys = ("mf", "sf", "nisf", "sfTP")
    print(json.dumps({"dataset": name, "n": int(len(d)), "s": round(time.time() - t0, 1),
                      "rmse": {k: rm[k] for k in keys}, "rmse_published": dict(zip(keys, pub[:4].tolist())),
                      "rmse_rel_dev": {k: abs(rm[k] - pub[i]) / pub[i] for i, k in enumerate(keys)},
                      "wrmse": {k: float(wm[k]) for k in keys}, "wrmse_published": dict(zip(keys, pub[4:].tolist()))}), flush=True)
