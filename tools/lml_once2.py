import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from skopt_b200.engine import DeviceFit
n, d = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.RandomState(0)
X = rng.rand(n, d); y = np.sin(X.dot(rng.randn(d))) + 0.1 * rng.randn(n); y = (y - y.mean()) / y.std()
fit = DeviceFit(X, y, 3, 1e-10)
theta = np.concatenate([[0.0], np.zeros(d), [np.log(1e-2)]])
fit.lml(theta, eval_gradient=True)
fit.lml(theta, eval_gradient=True)
fit.close()
