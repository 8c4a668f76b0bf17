"""Shared helpers: golden-case loading and score comparison (test infrastructure)."""
import json
import os

import numpy as np
import torch

from oracle import ptq_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLD, "cases.json")) as f:
    CASES = json.load(f)
COMMON = CASES["common"]


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    scores = [z[k] for k in sorted(k for k in z.files if k.startswith("scores_"))]
    return z, scores


def linear_case(name):
    case = CASES["linear"][name]
    fx = O.make_linear_fixture(**case["fx"])
    mod = case["mod"]
    sp = O.LinearSpec(fx[1].shape[1], fx[1].shape[0], n_V=mod["n_V"], n_H=mod["n_H"], n_a=mod["n_a"],
                      w_bit=mod["w_bit"], a_bit=mod["a_bit"], eq_alpha=COMMON["eq_alpha"], eq_beta=COMMON["eq_beta"],
                      eq_n=COMMON["eq_n"], search_round=mod["search_round"], post_gelu=bool(case.get("post_gelu")))
    return sp, fx, case


def matmul_case(name):
    case = CASES["matmul"][name]
    fx = O.make_matmul_fixture(**case["fx"])
    mod = case["mod"]
    sp = O.MatMulSpec(A_bit=mod["A_bit"], B_bit=mod["B_bit"], eq_alpha=COMMON["eq_alpha"], eq_beta=COMMON["eq_beta"],
                      eq_n=COMMON["eq_n"], search_round=mod["search_round"], sos=bool(case.get("sos")))
    return sp, fx, case


def flatten_linear_log(log):
    """[(sw_list, sa_list), ...] -> list of score tables in the reference's argmax call order."""
    out = []
    for sw, sa in log:
        out.extend(sw)
        out.extend(sa)
    return out


def flatten_matmul_log(log):
    out = []
    for s1, s2 in log:
        out.append(s1)
        out.append(s2)
    return out


def assert_scores_close(got, ref, rtol, what=""):
    got = np.asarray(got, dtype=np.float64).reshape(np.asarray(ref).shape)
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref).max()
    err = np.abs(got - ref).max() / (scale + 1e-300)
    assert err < rtol, f"{what}: score mismatch rel-to-max {err:.3e} >= {rtol}"


def assert_choice_consistent(got_idx, ref_scores, eps, what=""):
    """The chosen candidate may differ from the oracle's argmax only when the
    oracle's own score gap between the two candidates is below eps (relative to
    the spread of the score column) -- near-ties are legitimately order dependent."""
    ref_scores = np.asarray(ref_scores, dtype=np.float64)
    if ref_scores.ndim == 1:
        ref_scores = ref_scores[:, None]
    got_idx = np.asarray(got_idx).reshape(-1)
    best = ref_scores.argmax(0)
    for j in range(ref_scores.shape[1]):
        if got_idx[j] != best[j]:
            col = ref_scores[:, j]
            gap = col[best[j]] - col[got_idx[j]]
            spread = np.abs(col[best[j]]) + 1e-300
            assert gap / spread < eps, f"{what}: group {j} chose {got_idx[j]} vs oracle {best[j]}, gap {gap/spread:.3e}"


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))
