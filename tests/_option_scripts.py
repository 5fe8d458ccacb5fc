"""Driver functions shared by oracle/make_golden_traj_options.py (run against the unmodified reference to produce
tests/golden/traj_options.json) and tests/test_host_optimizer.py (run against this package): the SAME configurations and
the same Optimizer script, with the package under test passed in as `mod`.  TEST INFRASTRUCTURE ONLY."""
import numpy as np

OBJECTIVES = {
    "smooth2": lambda x: (x[0] - 0.3) ** 2 + np.sin(3 * x[1]) + 0.05 * x[0] * x[1],
    "mixed3": lambda x: (x[0] - 0.5) ** 2 + 0.3 * (x[1] - 3) ** 2 + {"a": 0.0, "b": 1.0, "c": -0.5}[x[2]],
    "cats2": lambda x: {"a": 0.0, "b": 1.0, "c": -0.5}[x[0]] + {"u": 0.3, "v": -0.2}[x[1]],
    "logs2": lambda x: np.log10(x[0]) ** 2 + 0.01 * x[1],
    "ints2": lambda x: (x[0] - 3) ** 2 + abs(x[1]),
}
R2 = [(-1.0, 2.0), (0.0, 3.0)]
BASE = dict(acq_optimizer="sampling", n_points=300)
RUNS = {
    "xi_large": dict(func="smooth2", dims=R2, acq_func="EI", xi=0.1, n_calls=9, n_initial_points=4, random_state=3),
    "pi_xi_small": dict(func="smooth2", dims=R2, acq_func="PI", xi=0.001, n_calls=9, n_initial_points=4, random_state=3),
    "kappa_small": dict(func="smooth2", dims=R2, acq_func="LCB", kappa=0.5, n_calls=9, n_initial_points=4, random_state=3),
    "kappa_large": dict(func="smooth2", dims=R2, acq_func="LCB", kappa=10.0, n_calls=9, n_initial_points=4, random_state=3),
    "x0_y0": dict(func="smooth2", dims=R2, acq_func="EI", n_calls=7, n_initial_points=2, random_state=1,
                  x0=[[0.1, 0.2], [1.5, 2.5]], y0="evaluate"),
    "x0_single": dict(func="smooth2", dims=R2, acq_func="EI", n_calls=7, n_initial_points=2, random_state=1, x0=[0.1, 0.2]),
    "no_random_phase": dict(func="smooth2", dims=R2, acq_func="LCB", n_calls=3, n_initial_points=0, random_state=1,
                            x0=[[0.1, 0.2], [1.5, 2.5], [0.7, 0.1]]),
    "noise_1e-3": dict(func="smooth2", dims=R2, acq_func="EI", n_calls=8, n_initial_points=4, random_state=5, noise=1e-3),
    "queue_2": dict(func="smooth2", dims=R2, acq_func="EI", n_calls=8, n_initial_points=3, random_state=5,
                    model_queue_size=2),
    "mixed_hedge": dict(func="mixed3", dims=[(0.0, 1.0), (1, 5), ["a", "b", "c"]], acq_func="gp_hedge", n_calls=9,
                        n_initial_points=4, random_state=7),
    "categorical_only": dict(func="cats2", dims=[["a", "b", "c"], ["u", "v"]], acq_func="EI", n_calls=8,
                             n_initial_points=3, random_state=2, acq_optimizer="auto", n_points=100),
    "log_uniform": dict(func="logs2", dims=[(1e-3, 1e2, "log-uniform"), (1, 100, "log-uniform")], acq_func="EI",
                        n_calls=8, n_initial_points=4, random_state=4),
    "integers_only": dict(func="ints2", dims=[(0, 10), (-5, 5)], acq_func="PI", n_calls=8, n_initial_points=4,
                          random_state=4),
}
SCRIPT_CASES = [["EI", {}], ["gp_hedge", {}], ["LCB", {"acq_func_kwargs": {"kappa": 3.0}}], ["EIps", {}], ["PIps", {}],
                ["gp_hedge", {"acq_func_kwargs": {"eta": 0.1}}], ["EI", {"model_queue_size": 1}]]


def plain(o):
    return o.item() if hasattr(o, "item") else str(o)


def run_config(mod, cfg):
    kw = dict(BASE)
    kw.update({k: v for k, v in cfg.items() if k not in ("func", "dims")})
    func = OBJECTIVES[cfg["func"]]
    dims = [tuple(d) if not isinstance(d[0], str) else list(d) for d in cfg["dims"]]
    if kw.get("y0") == "evaluate":
        kw["y0"] = [func(x) for x in kw["x0"]]
    res = mod.gp_minimize(func, dims, **kw)
    return {"x_iters": res.x_iters, "func_vals": list(res.func_vals), "x": res.x, "n_models": len(res.models)}


def run_script(mod, acq, okw):
    """One Optimizer driven through its whole public surface; every observable is logged."""
    func = OBJECTIVES["smooth2"]
    log = []
    opt = mod.Optimizer(R2, "GP", n_initial_points=3, acq_func=acq, acq_optimizer="sampling",
                        acq_optimizer_kwargs=dict(n_points=150), random_state=9, **okw)
    per_second = "ps" in acq
    value = (lambda x: (func(x), 1.0 + abs(x[0]))) if per_second else func
    xs = opt.ask(n_points=3)
    log.append(["ask3", xs])
    opt.tell(xs, [value(x) for x in xs])
    log.append(["models", len(opt.models)])
    x = opt.ask()
    log.append(["ask", x, opt.ask()])
    opt.tell(x, value(x), fit=False)
    log.append(["nofit-next", opt.ask()])
    x = [0.5, 0.5]
    res = opt.tell(x, value(x))
    log.append(["told", opt.ask(), len(res.models), float(res.fun)])
    log.append(["constant-liar", opt.ask(n_points=2, strategy="cl_mean"), opt.ask(n_points=2, strategy="cl_max")])
    opt.update_next()
    log.append(["after-update_next", opt.ask()])
    if not per_second:
        res = opt.run(func, n_iter=2)
        log.append(["run", res.x_iters[-2:], res.x])
    twin = opt.copy(random_state=4)
    log.append(["copy", twin.ask(), len(twin.models)])
    res = opt.get_result()
    log.append(["result", res.x, len(res.x_iters), sorted(res.keys())])
    return log
