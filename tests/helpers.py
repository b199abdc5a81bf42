"""Shared helpers for the tests (oracle side only)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["burgers_w8_d2", "burgers_w20_d8", "helmholtz_w16_d3", "allencahn_w16_d2",
                "heat_w8_d3", "navierstokes_w16_d3", "helmholtz_w64_d4"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = json.loads(str(z["cfg_json"]))
    return cfg, z


def cfg_from_dict(O, d, **kw):
    c = O.make_config(d["pde"], d["width"], d["depth"], d["n_interior"], d["n_boundary"], d["n_initial"])
    for k in ("seed_interior", "seed_boundary", "seed_initial", "seed_weights", "lr", "beta1", "beta2", "eps",
              "lambda_r", "lambda_b", "lambda_0", "pde_param"):
        setattr(c, k, d[k])
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
