#!/usr/bin/env python
"""Golden vectors for the year-block resampling of the bootstrap, made by EXECUTING the reference's own code.

``get_random_block_indices`` is a nested function of ``fit_and_detrend_cell`` (attrici/detrend.py:500-531) and
cannot be imported; this script cuts its source text out of /root/reference/attrici/detrend.py (unmodified, only
dedented), executes it with ``blocks`` / ``np`` in scope and records, for several calendars and seeds, the index
vector the reference builds at detrend.py:541-546.

    python tests/golden/make_golden_bootstrap.py   ->   tests/golden/bootstrap_blocks.npz
"""
import os
import re
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = "/root/reference/attrici/detrend.py"


def reference_function(blocks):
    text = open(SRC).read()
    m = re.search(r"^( +)def get_random_block_indices\(target_length\):\n(?:.*\n)*?\1    return block_indices\[:target_length\]\n",
                  text, re.M)
    ns = {"np": np, "blocks": blocks}
    exec(textwrap.dedent(m.group(0)), ns)
    return ns["get_random_block_indices"]


def calendar(start, stop):
    t = np.arange(np.datetime64(start), np.datetime64(stop) + 1)
    return t.astype("datetime64[Y]").astype(int) + 1970


def main():
    out = {}
    cases = [("1901-01-01", "2019-12-31", 11, 3), ("1979-01-01", "2000-12-31", 5, 4), ("1950-03-10", "1960-09-30", 7, 3)]
    for i, (a, b, seed, n) in enumerate(cases):
        years = calendar(a, b)
        blocks = [np.flatnonzero(years == y) for y in np.unique(years)]   # detrend.py:491-498
        f = reference_function(blocks)
        np.random.seed(seed)
        idx = [np.concatenate([f(len(bl)) for bl in blocks]) for _ in range(n)]   # detrend.py:541-546
        out[f"case{i}_years"] = years.astype(np.int32)
        out[f"case{i}_seed"] = np.int64(seed)
        out[f"case{i}_indices"] = np.asarray(idx, dtype=np.int32)
    out["n_cases"] = np.int64(len(cases))
    path = os.path.join(ROOT, "tests", "golden", "bootstrap_blocks.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
