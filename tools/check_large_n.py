"""Ad-hoc sanity check at a large, ragged n_train (index arithmetic, padded blocks): FP64 vs split vs oracle sample."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skopt_b200
from skopt_b200 import engine
from oracle import gp_oracle as orc
from tests.test_gpu_parity import host_state_from_oracle_state

n, d, m = int(sys.argv[1]) if len(sys.argv) > 1 else 6001, 8, 20000
t0 = time.time()
st = orc.make_synthetic_state(n, d, noise=1e-3, seed=4)
print("state built in %.1f s" % (time.time() - t0))
Xc = np.random.RandomState(1).rand(m, d)
dev = engine.DeviceState(host_state_from_oracle_state(st))
a = dev.score(Xc, y_opt=0.0, acq_funcs=("EI", "LCB"), top_k=5, precision="fp64")
b = dev.score(Xc, y_opt=0.0, acq_funcs=("EI", "LCB"), top_k=5, precision="split_int8")
g = dev.score_grad(Xc[:300], y_opt=0.0, acq_funcs=("EI",), precision="fp64")
gs = dev.score_grad(Xc[:300], y_opt=0.0, acq_funcs=("EI",), precision="split_int8")
idx = np.arange(0, m, m // 200)
mu, sd = orc.predict(st, Xc[idx], return_std=True)
scale = st.prior_var * st.y_std ** 2
print("fp64  vs oracle: mean %.2e  var %.2e" % (np.max(np.abs(a["mean"][idx] - mu)) / st.y_std, np.max(np.abs(a["std"][idx] ** 2 - sd ** 2)) / scale))
print("split vs oracle: mean %.2e  var %.2e" % (np.max(np.abs(b["mean"][idx] - mu)) / st.y_std, np.max(np.abs(b["std"][idx] ** 2 - sd ** 2)) / scale))
print("split vs fp64  : var %.2e  argmin equal %s" % (np.max(np.abs(a["std"] ** 2 - b["std"] ** 2)) / scale, a["EI"]["topk_idx"][0] == b["EI"]["topk_idx"][0]))
print("grad fp64 vs split: mean_grad %.2e std_grad*std %.2e; grad-path std vs value-path %.2e" % (
    np.max(np.abs(g["mean_grad"] - gs["mean_grad"])) / np.max(np.abs(g["mean_grad"])),
    np.max(np.abs((g["std_grad"] - gs["std_grad"]) * g["std"][:, None])) / scale,
    np.max(np.abs(g["std"] ** 2 - a["std"][:300] ** 2)) / scale))
assert np.max(np.abs(a["std"][idx] ** 2 - sd ** 2)) <= 1e-9 * scale and np.max(np.abs(b["std"][idx] ** 2 - sd ** 2)) <= 1e-9 * scale
print("OK")
