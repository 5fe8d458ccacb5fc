"""Live differential check of the Python boundary against the UNMODIFIED reference (build container only: reads
/root/reference).  The same gp_minimize configurations and the same Optimizer ask / tell script are driven through
`skopt` and through `skopt_b200`; every asked point, function value and result field must be equal.  Without a GPU the
host entry points of the C ABI are the software model tests/_abi_model.py (TEST INFRASTRUCTURE; arithmetic by the CPU
oracle), so this exercises everything above the ABI; pass --device on a box with a B200 and the reference tree to use the
real libskb.so.  Sampling-mode runs must be identical; L-BFGS runs ("auto" on spaces with gradients) agree to the last
digits only and are reported, not counted.

    python tools/diff_against_reference.py [--device]
"""
import json
import os
import sys
import warnings

import numpy as np

if not hasattr(np, "int"):
    np.int = int                      # the reference predates NumPy 1.24
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(1, "/root/reference")
warnings.simplefilter("ignore")
import skopt as R  # noqa: E402
import skopt_b200 as O  # noqa: E402

if "--device" not in sys.argv:
    from skopt_b200 import _lib  # noqa: E402
    from skopt_b200.optimizer import optimizer as om  # noqa: E402
    from tests._abi_model import AbiModel  # noqa: E402
    _lib._lib = AbiModel()
    om.device_dim_descs = lambda s: None

# ---- gp_minimize configurations ----
def f2(x): return (x[0] - 0.3) ** 2 + np.sin(3 * x[1]) + 0.05 * x[0] * x[1]
def fm(x): return (x[0] - 0.5) ** 2 + 0.3 * (x[1] - 3) ** 2 + {"a": 0.0, "b": 1.0, "c": -0.5}[x[2]]
def fc(x): return {"a": 0.0, "b": 1.0, "c": -0.5}[x[0]] + {"u": 0.3, "v": -0.2}[x[1]]
cases = []
for acq, kw in [("EI", dict(xi=0.1)), ("PI", dict(xi=0.001)), ("LCB", dict(kappa=0.5)), ("gp_hedge", {}), ("LCB", dict(kappa=10.0))]:
    cases.append(("real2-" + acq + json.dumps(kw), f2, [(-1.0, 2.0), (0.0, 3.0)], dict(acq_func=acq, acq_optimizer="sampling", n_calls=9, n_initial_points=4, n_points=300, random_state=3, **kw)))
cases.append(("x0y0", f2, [(-1.0, 2.0), (0.0, 3.0)], dict(acq_func="EI", acq_optimizer="sampling", n_calls=7, n_initial_points=2, n_points=200, random_state=1, x0=[[0.1, 0.2], [1.5, 2.5]], y0=[f2([0.1, 0.2]), f2([1.5, 2.5])])))
cases.append(("x0only", f2, [(-1.0, 2.0), (0.0, 3.0)], dict(acq_func="EI", acq_optimizer="sampling", n_calls=7, n_initial_points=2, n_points=200, random_state=1, x0=[0.1, 0.2])))
cases.append(("noinit", f2, [(-1.0, 2.0), (0.0, 3.0)], dict(acq_func="LCB", acq_optimizer="sampling", n_calls=3, n_initial_points=0, n_points=200, random_state=1, x0=[[0.1, 0.2], [1.5, 2.5], [0.7, 0.1]])))
cases.append(("noise1e-3", f2, [(-1.0, 2.0), (0.0, 3.0)], dict(acq_func="EI", acq_optimizer="sampling", n_calls=8, n_initial_points=4, n_points=200, random_state=5, noise=1e-3)))
cases.append(("queue2", f2, [(-1.0, 2.0), (0.0, 3.0)], dict(acq_func="EI", acq_optimizer="sampling", n_calls=8, n_initial_points=3, n_points=200, random_state=5, model_queue_size=2)))
cases.append(("mixed", fm, [(0.0, 1.0), (1, 5), ["a", "b", "c"]], dict(acq_func="gp_hedge", acq_optimizer="sampling", n_calls=9, n_initial_points=4, n_points=300, random_state=7)))
cases.append(("mixed-auto", fm, [(0.0, 1.0), (1, 5), ["a", "b", "c"]], dict(acq_func="EI", acq_optimizer="auto", n_calls=7, n_initial_points=4, n_points=300, random_state=7)))
cases.append(("cat", fc, [["a", "b", "c"], ["u", "v"]], dict(acq_func="EI", n_calls=8, n_initial_points=3, n_points=100, random_state=2)))
cases.append(("logunif", lambda x: np.log10(x[0]) ** 2 + x[1], [(1e-3, 1e2, "log-uniform"), (1, 100, "log-uniform")], dict(acq_func="EI", acq_optimizer="sampling", n_calls=8, n_initial_points=4, n_points=200, random_state=4)))
cases.append(("int-only", lambda x: (x[0] - 3) ** 2 + abs(x[1]), [(0, 10), (-5, 5)], dict(acq_func="PI", acq_optimizer="sampling", n_calls=8, n_initial_points=4, n_points=200, random_state=4)))
bad = 0
for name, func, dims, kw in cases:
    try:
        a = R.gp_minimize(func, dims, **kw)
    except Exception as e:
        a = e
    try:
        b = O.gp_minimize(func, dims, **kw)
    except Exception as e:
        b = e
    if isinstance(a, Exception) or isinstance(b, Exception):
        same = type(a) == type(b)
        print(name, "EXC", repr(a)[:100], "|", repr(b)[:100], "OK" if same else "MISMATCH"); bad += not same
        continue
    same = a.x_iters == b.x_iters and np.array_equal(a.func_vals, b.func_vals) and a.x == b.x and len(a.models) == len(b.models)
    k = next((i for i, (p, q) in enumerate(zip(a.x_iters, b.x_iters)) if p != q), None)
    refined = kw.get("acq_optimizer") == "auto" and not isinstance(dims[0], list)
    print("%-40s %s" % (name, "identical" if same else "%s at call %s: %s vs %s" % (
        "L-BFGS last digits" if refined else "DIFFERS", k, a.x_iters[k] if k is not None else None,
        b.x_iters[k] if k is not None else None)))
    bad += (not same) and not refined

# ---- one Optimizer ask / tell script per acquisition ----
def f(x): return (x[0] - 0.3) ** 2 + np.sin(3 * x[1])
dims = [(-1.0, 2.0), (0.0, 3.0)]
def script(mod, acq="EI", **okw):
    log = []
    opt = mod.Optimizer(dims, "GP", n_initial_points=3, acq_func=acq, acq_optimizer="sampling",
                        acq_optimizer_kwargs=dict(n_points=150), random_state=9, **okw)
    ps = "ps" in acq
    val = (lambda x: (f(x), 1.0 + abs(x[0]))) if ps else f
    xs = opt.ask(n_points=3); log.append(("ask3", xs))                       # batch during the random phase
    opt.tell(xs, [val(x) for x in xs]); log.append(("models", len(opt.models)))
    x = opt.ask(); log.append(("ask", x)); log.append(("ask-again", opt.ask()))
    opt.tell(x, val(x), fit=False); log.append(("nofit-next", opt.ask()))
    x = [0.5, 0.5]; res = opt.tell(x, val(x)); log.append(("told", opt.ask(), len(res.models), float(res.fun)))
    log.append(("cl", opt.ask(n_points=2, strategy="cl_mean"), opt.ask(n_points=2, strategy="cl_max")))
    opt.update_next(); log.append(("after-update_next", opt.ask()))
    if not ps:
        res = opt.run(f, n_iter=2); log.append(("run", res.x_iters[-2:], res.x))
    c = opt.copy(random_state=4); log.append(("copy", c.ask(), len(c.models)))
    r = opt.get_result(); log.append(("result", r.x, len(r.x_iters), sorted(r.keys())))
    return log
for acq, okw in [("EI", {}), ("gp_hedge", {}), ("LCB", dict(acq_func_kwargs=dict(kappa=3.0))), ("EIps", {}), ("PIps", {}),
                 ("gp_hedge", dict(acq_func_kwargs=dict(eta=0.1))), ("EI", dict(model_queue_size=1))]:
    a, b = script(R, acq, **okw), script(O, acq, **okw)
    for (ka, *va), (kb, *vb) in zip(a, b):
        sa, sb = json.dumps(va, default=lambda o: o.item() if hasattr(o, "item") else str(o)), json.dumps(vb, default=lambda o: o.item() if hasattr(o, "item") else str(o))
        if sa != sb:
            bad += 1; print(acq, okw, ka, "\n   ref:", sa[:300], "\n   our:", sb[:300])
    print(acq, okw, "steps compared:", len(a))
print("mismatches:", bad)
sys.exit(1 if bad else 0)
