"""Shared helpers of the test-suite: golden fixtures (tests/golden/*.npz, generated from the unmodified
reference by tests/golden/make_golden.py) and the oracle-side preparation of a fixture cell."""

import glob
import os

import numpy as np

from oracle import attrici_oracle as oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
MODES = 4

FIXTURES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*_lat*.npz")))
NORMAL_FIXTURES = [f for f in FIXTURES if f.split("_")[0] in ("tas", "ps", "rlds", "rsds", "tasskew")]
ONE_PER_VARIABLE = ["tas_lat50.75_lon9.25"] + [f for f in FIXTURES if not f.startswith("tas_")]


def load(name):
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    d["variable"] = str(d["variable"])
    d["lat"], d["lon"] = float(d["lat"]), float(d["lon"])
    d["n_fit"] = int(d["n_fit"])
    return d


def predictor(n):
    g = np.load(os.path.join(GOLDEN, "ssa_gmt.npz"))
    return oracle.gmt_on_time_axis(g["tas"], n)


def replaced_codes(rep):
    """reference `replaced` (NaN / +Inf / -Inf / 0 as float64) -> library codes 1 / 2 / 3 / 0"""
    return np.where(np.isnan(rep), 1, np.where(rep == np.inf, 2, np.where(rep == -np.inf, 3, 0))).astype(np.int32)


class Cell:
    """Everything the oracle derives for one fixture cell before fitting."""

    def __init__(self, name):
        d = load(name)
        self.d = d
        self.name = name
        self.variable = d["variable"]
        self.family, self.kind, self.lo, self.hi, self.post = oracle.VARIABLES[self.variable]
        self.n = len(d["y"])
        self.days = np.arange(self.n, dtype=np.int32)
        self.gmt = predictor(self.n)
        self.ys, self.scaling, self.y_out = oracle.scale_variable(self.variable, d["y"])
        self.n_fit = d["n_fit"]
        self.problem = oracle.build_problem(self.family, self.ys, self.gmt, self.days, MODES, slice(0, self.n_fit))
        self.uniforms = None
        if self.family == "bernoulli_gamma":
            np.random.seed(oracle.cell_seed(0, d["lat"], d["lon"]))
            self.uniforms = np.random.rand(self.n)
