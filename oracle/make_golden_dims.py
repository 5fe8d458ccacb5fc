"""Fixtures for the search-space API either side of the hot path: what every Dimension and
transformer of the UNMODIFIED reference returns - values AND container / scalar types - for sampling, warping and
un-warping, so that the array-first restatement in skopt_b200.space can be pinned without the reference on the box.
Run in the build container only.  TEST INFRASTRUCTURE ONLY."""
import json
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
if not hasattr(np, "int"):
    np.int = int           # the reference predates NumPy 1.24 (transformers.py:262,275)
import skopt.space as ref  # noqa: E402
from skopt.space import transformers as ref_t  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "space_api.json")

DIMENSIONS = [
    ["Real", dict(low=-2.5, high=7.0)],
    ["Real", dict(low=1e-3, high=1e2, prior="log-uniform")],
    ["Real", dict(low=1.0, high=64.0, prior="log-uniform", base=2)],
    ["Real", dict(low=0.0, high=1.0, dtype="float32")],
    ["Integer", dict(low=-3, high=11)],
    ["Integer", dict(low=1, high=1000, prior="log-uniform")],
    ["Integer", dict(low=1, high=64, prior="log-uniform", base=2)],
    ["Integer", dict(low=0, high=9, dtype="int32")],
    ["Integer", dict(low=0, high=9, dtype="int")],
    ["Categorical", dict(categories=["b", "a", "c"])],
    ["Categorical", dict(categories=["x", "y"])],
    ["Categorical", dict(categories=[3, 1, 2, 10])],
    ["Categorical", dict(categories=[True, False])],
    ["Categorical", dict(categories=["only"])],
    ["Categorical", dict(categories=["p", "q", "r"], prior=[0.2, 0.5, 0.3])],
    ["Categorical", dict(categories=list("abcdefghij"))],
]


def tagged(v):
    """JSON form that keeps what a caller can observe: container kind, element type names, dtype, shape, values."""
    if isinstance(v, np.ndarray):
        return {"kind": "ndarray", "dtype": v.dtype.kind, "shape": list(v.shape), "values": v.tolist()}
    if isinstance(v, (list, tuple)):
        return {"kind": type(v).__name__, "items": [tagged(x) for x in v]}
    if isinstance(v, np.generic):
        return {"kind": "npscalar", "dtype": v.dtype.kind, "values": v.item()}
    return {"kind": type(v).__name__, "values": v}


def attempt(fn):
    try:
        return tagged(fn())
    except Exception as exc:                       # the error class is part of the contract
        return {"kind": "raises", "values": type(exc).__name__}


def dimension_records():
    out = []
    for cls, kwargs in DIMENSIONS:
        transforms = ["onehot", "label", "string", "identity", "normalize"] if cls == "Categorical" else \
            ["identity", "normalize"]
        for transform in transforms:
            dim = getattr(ref, cls)(transform=transform, **kwargs)
            rec = {"cls": cls, "kwargs": kwargs, "transform": transform, "repr": repr(dim),
                   "transformed_size": dim.transformed_size, "bounds": tagged(dim.bounds),
                   "transformed_bounds": tagged(dim.transformed_bounds), "is_constant": bool(dim.is_constant),
                   "rvs": []}
            for n in ([None, 1, 7] if cls == "Categorical" else [1, 7]):
                for seed in (0, 3):
                    rec["rvs"].append([n, seed, attempt(lambda: dim.rvs(n_samples=n, random_state=seed))])
            points = dim.rvs(n_samples=9, random_state=5)
            rec["points"] = tagged(points)
            warped = dim.transform(points)
            rec["transform_out"] = tagged(np.asarray(warped)) if transform not in ("string", "identity") or \
                cls != "Categorical" else tagged(list(warped))
            rec["inverse_out"] = attempt(lambda: dim.inverse_transform(warped))
            if cls != "Categorical":
                first = float(np.asarray(warped, dtype=float).ravel()[0])
                rec["inverse_scalar"] = [first, attempt(lambda: dim.inverse_transform(first))]
            out.append(rec)
    return out


def transformer_records():
    rng = np.random.RandomState(11)
    x = rng.uniform(-3.0, 5.0, 13)
    recs = {}
    norm = ref_t.Normalize(-3.0, 5.0)
    recs["normalize"] = {"low": -3.0, "high": 5.0, "x": x.tolist(), "t": norm.transform(x).tolist(),
                         "back": norm.inverse_transform(norm.transform(x)).tolist(),
                         "too_high": attempt(lambda: norm.transform([5.0 + 1e-7])),
                         "too_low": attempt(lambda: norm.transform([-3.0 - 1e-7])),
                         "within_slack": tagged(norm.transform([5.0 + 1e-9])),
                         "inverse_too_high": attempt(lambda: norm.inverse_transform([1.0 + 1e-7]))}
    inorm = ref_t.Normalize(1, 20, is_int=True)
    xi = rng.randint(1, 21, 11)
    recs["normalize_int"] = {"low": 1, "high": 20, "x": xi.tolist(), "t": inorm.transform(xi).tolist(),
                             "back": tagged(inorm.inverse_transform(inorm.transform(xi))),
                             "round_in": tagged(inorm.transform([19.8, 1.3])),
                             "too_high": attempt(lambda: inorm.transform([20.6])),
                             "degenerate": tagged(ref_t.Normalize(4, 4, is_int=True).transform([4, 4]))}
    for base in (2, 10):
        logn = ref_t.LogN(base)
        xp = rng.uniform(0.01, 50.0, 9)
        recs["logn_%d" % base] = {"x": xp.tolist(), "t": logn.transform(xp).tolist(),
                                  "back": logn.inverse_transform(logn.transform(xp)).tolist()}
    for name, cats in (("str", ["b", "a", "c"]), ("int", [3, 1, 2, 10]), ("two", ["x", "y"]), ("one", ["only"])):
        picks = [cats[i] for i in rng.randint(0, len(cats), 8)]
        lab = ref_t.LabelEncoder(cats)
        hot = ref_t.CategoricalEncoder().fit(cats)
        st = ref_t.StringEncoder()
        st.fit(cats)
        recs["encoders_" + name] = {
            "categories": cats, "picks": picks,
            "label": tagged(lab.transform(picks)), "label_back": tagged(lab.inverse_transform(lab.transform(picks))),
            "label_back_float": tagged(lab.inverse_transform(np.asarray(lab.transform(picks), dtype=float) + 0.3)),
            "onehot": tagged(hot.transform(picks)), "onehot_back": tagged(hot.inverse_transform(hot.transform(picks))),
            "string": tagged(st.transform(picks)), "string_back": tagged(st.inverse_transform(st.transform(picks)))}
    pipe = ref_t.Pipeline([ref_t.LogN(10), ref_t.Normalize(-2.0, 2.0)])
    xp = rng.uniform(0.01, 100.0, 7)
    recs["pipeline"] = {"x": xp.tolist(), "t": pipe.transform(xp).tolist(),
                        "back": pipe.inverse_transform(pipe.transform(xp)).tolist(),
                        "bad_member": attempt(lambda: ref_t.Pipeline([ref_t.LogN(10), "normalize"]))}
    return recs


def main():
    rec = {"dimensions": dimension_records(), "transformers": transformer_records(),
           "versions": {"numpy": np.__version__}}
    with open(OUT, "w") as f:
        json.dump(rec, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
