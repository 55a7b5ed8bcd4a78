"""Clocks and board power of the two kernels of the INT8 posterior pipeline, each running ALONE for a few seconds
(GPC_I8_PHASE=kstar | vt, a profiling switch inside gpc_predict_dev) and together: is the step bound by the power cap?
    python profiles/tools/phase_power.py [seconds]
"""
import json, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as entry
entry.setup_path()
import torch, gpcore, bench
from gpcore import _lib as L

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
N, M, F = 2048, 1000000, 2
X4, y = bench.make_train(N, F)
core = gpcore.GPCore(L.KIND_MF_AR1_RBF, F, 0)
core.set_hypers(bench.MF2_PARAMS, 1e-8); core.set_data(X4, y); core.factor()
dXs = torch.from_numpy(bench.make_grid(100, 1)).cuda()
dm = torch.empty(M, dtype=torch.float64, device="cuda"); dv = torch.empty_like(dm)
flags = L.INCLUDE_NOISE | L.CLIP_DIAG
stream = torch.cuda.ExternalStream(core.stream())


class Smi:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                                   "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._r, daemon=True).start()
    def _r(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))
    def window(self, t0, t1):
        r = [x for t, x in self.rows if t0 <= t <= t1]
        clk = [float(x[0]) for x in r]; pw = [float(x[1]) for x in r]
        return {"sm_mhz_median": float(np.median(clk)) if clk else None, "power_w_median": float(np.median(pw)) if pw else None,
                "power_w_max": max(pw) if pw else None, "power_cap_active": sum(x[2].lower().startswith("active") for x in r), "samples": len(r)}

smi = Smi()
out = {}
for phase in ("all", "kstar", "vt"):
    if phase == "all":
        os.environ.pop("GPC_I8_PHASE", None)
    else:
        os.environ["GPC_I8_PHASE"] = phase
    for _ in range(3):
        core.predict_dev(dXs.data_ptr(), M, dm.data_ptr(), dv.data_ptr(), flags)
    torch.cuda.synchronize()
    time.sleep(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); steps = 0
    e0.record(stream)
    while time.time() - t0 < secs:
        for _ in range(5):
            core.predict_dev(dXs.data_ptr(), M, dm.data_ptr(), dv.data_ptr(), flags)
        steps += 5
        torch.cuda.synchronize()
    e1.record(stream); torch.cuda.synchronize()
    t1 = time.time()
    out[phase] = dict(ms_per_step=e0.elapsed_time(e1) / steps, steps=steps, **smi.window(t0 + 0.5 * secs, t1))
    time.sleep(1.0)
smi.p.terminate()
print(json.dumps(out))
