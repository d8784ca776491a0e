"""Import shim for the upstream reference (`pytens`) -- TEST INFRASTRUCTURE ONLY.

The reference lives read-only at /root/reference and cannot be imported
verbatim in the build container: `opt_einsum`, `matplotlib`, `line_profiler`,
`tntorch` and `gurobipy` are absent.  This module registers
stub modules for those five names so that `import pytens` succeeds, and exposes
`load_reference()`.

It exists only to (a) validate the numpy restatement in `oracle/tt_oracle.py`
and (b) generate the golden fixtures under `tests/golden/` (see
`oracle/make_golden.py`).  /root/reference does not exist on the GPU box, so
nothing on the product path, in the `-m gpu` tests, `smoke()` or `bench.py`
may import this file.

The `opt_einsum.contract` stand-in evaluates the einsum with numpy's own
path optimiser (`np.einsum(..., optimize="greedy")`) using integer sublists,
so it is not limited to 26 letters (pytens maps index i -> chr(97+i),
pytens/algs.py:454-456, which runs past 'z' for a TT pair with d >= 10).
"""

from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PYTENS_REFERENCE_ROOT", "/root/reference")


def _contract(estr, *arrs, optimize="auto", **_kw):
    """Stand-in for opt_einsum.contract (call site pytens/algs.py:482, :1172)."""
    lhs, rhs = estr.split("->")
    terms = lhs.split(",")
    symbols = {}
    for t in terms:
        for ch in t:
            symbols.setdefault(ch, len(symbols))
    if len(symbols) > 52:
        raise ValueError("shim einsum supports at most 52 distinct indices")
    args = []
    for t, a in zip(terms, arrs):
        args.append(np.asarray(a))
        args.append([symbols[ch] for ch in t])
    args.append([symbols[ch] for ch in rhs])
    return np.einsum(*args, optimize="greedy")


def _install_stubs() -> None:
    if "opt_einsum" not in sys.modules:
        oe = types.ModuleType("opt_einsum")
        oe.contract = _contract
        sys.modules["opt_einsum"] = oe
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "line_profiler" not in sys.modules:
        lp = types.ModuleType("line_profiler")
        lp.profile = lambda f: f
        sys.modules["line_profiler"] = lp
    if "tntorch" not in sys.modules:
        tn = types.ModuleType("tntorch")
        mv = types.ModuleType("tntorch.maxvol")

        def py_maxvol(*_a, **_k):
            raise RuntimeError("tntorch is not available (shim)")

        mv.py_maxvol = py_maxvol
        tn.maxvol = mv
        sys.modules["tntorch"] = tn
        sys.modules["tntorch.maxvol"] = mv
    if "gurobipy" not in sys.modules:
        gp = types.ModuleType("gurobipy")
        gp.GRB = types.SimpleNamespace()
        sys.modules["gurobipy"] = gp


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pytens"))


def load_reference():
    """Return the imported reference package `pytens` (shimmed)."""
    if not reference_available():
        raise ImportError(f"reference not present at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pytens  # noqa: E402  pylint: disable=import-error

    return pytens
