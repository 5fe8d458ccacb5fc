"""Break down the host-side cost of one model-based tell() at small sizes (where latency, not throughput, matters)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skopt_b200
from skopt_b200 import Optimizer, engine
from bench import branin

opt = Optimizer([(-5.0, 10.0), (0.0, 15.0)], "GP", n_initial_points=10, acq_func="gp_hedge", acq_optimizer="sampling",
                acq_optimizer_kwargs=dict(n_points=10000), random_state=0)
for _ in range(25):
    x = opt.ask(); opt.tell(x, branin(x))
est = opt.models[-1]
def t(f, n=20):
    f(); t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
print("rvs_transformed 10k x2      %.3f ms" % t(lambda: opt.space.rvs_transformed(10000, opt.rng)))
print("host_state_from_estimator   %.3f ms" % t(lambda: engine.host_state_from_estimator(est)))
hs = engine.host_state_from_estimator(est)
def mk():
    d = engine.DeviceState(hs); d.close()
print("DeviceState create+destroy  %.3f ms" % t(mk))
dev = engine.DeviceState(hs)
X = opt.space.rvs_transformed(10000, opt.rng)
print("score 3 acq top1            %.3f ms" % t(lambda: dev.score(X, y_opt=0.0, acq_funcs=("EI", "LCB", "PI"), top_k=1, want_posterior=False, want_values=False)))
print("predict_mean 3 pts          %.3f ms" % t(lambda: dev.predict_mean(X[:3])))
print("inverse_transform           %.3f ms" % t(lambda: opt.space.inverse_transform(X[:1])))
print("scoring_times median        %.3f ms" % (np.median(opt.scoring_times_[1:]) * 1e3))
