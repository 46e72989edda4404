"""CPU: anchoring of the restated geomloss solver (oracle/geomloss_ref.py -- PARITY UNPINNED, no wheel offline) on
optimal-transport theory, with code that shares nothing with the oracle or the kernels:

  * annealed slowly (scaling 0.99) the solver must land on the entropic-OT fixed point at eps = blur**p: its dual value
    must equal that of an independent Sinkhorn-Knopp iteration in scaling (not log) form, and its potentials must
    satisfy the two softmin equations;
  * for a small blur the value must approach the exact earth mover's distance from an independent linear program
    (scipy.optimize.linprog) -- this pins the conventions a from-memory restatement could get wrong: cost = distance
    (p=1, not squared, no 1/2), eps = blur**p, dual = <alpha,f> + <beta,g> without debiasing.

What stays unpinned is the truncated schedule itself (scaling 0.9: one averaged step per epsilon); that part follows the
published description of geomloss 0.2.4 and the reference's call sites (see the oracle's header).
"""
import numpy as np
import pytest
from scipy.optimize import linprog

from oracle import geomloss_ref as gr


def _problem(seed, n, m, d=16):
    rng = np.random.default_rng(seed)
    x, y = rng.normal(size=(n, d)), rng.normal(size=(m, d)) + 0.3
    a, b = rng.random(n) + 0.1, rng.random(m) + 0.1
    C = np.sqrt(((x[:, None, :] - y[None, :, :]) ** 2).sum(-1))
    return a / a.sum(), b / b.sum(), C


def _sinkhorn_knopp(a, b, C, eps, iters=20000):
    """Textbook matrix-scaling Sinkhorn; returns the dual value <a, f> + <b, g> of the entropic problem with the
    KL-to-product-measure regulariser (the convention geomloss uses: P = exp((f+g-C)/eps) a b)."""
    K = np.exp(-C / eps)
    u, v = np.ones_like(a), np.ones_like(b)
    for _ in range(iters):
        u = 1.0 / (K @ (b * v))
        v = 1.0 / (K.T @ (a * u))
    f, g = eps * np.log(u), eps * np.log(v)
    return float(a @ f + b @ g), f, g


def _emd(a, b, C):
    n, m = C.shape
    A_eq = np.zeros((n + m, n * m))
    for i in range(n):
        A_eq[i, i * m:(i + 1) * m] = 1
    for j in range(m):
        A_eq[n + j, j::m] = 1
    res = linprog(C.reshape(-1), A_eq=A_eq, b_eq=np.concatenate([a, b]), bounds=(0, None), method="highs")
    assert res.status == 0
    return float(res.fun)


@pytest.mark.parametrize("seed,n,m", [(0, 6, 9), (1, 10, 10), (2, 3, 12)])
def test_slowly_annealed_solver_reaches_the_entropic_fixed_point(seed, n, m):
    a, b, C = _problem(seed, n, m)
    blur = 0.5
    eps = gr.epsilon_schedule(1, float(C.max() * 2), blur, 0.99)
    f, g, dual = gr.sinkhorn_np(a[None], b[None], C[None], eps)
    want, f_sk, g_sk = _sinkhorn_knopp(a, b, C, blur)
    assert abs(dual[0] - want) <= 2e-4 * abs(want)
    # the two softmin equations of the fixed point (up to the constant shift f+c, g-c the dual is invariant to)
    f0, g0 = f[0], g[0]
    lhs_f = -blur * np.log((b[None, :] * np.exp((g0[None, :] - C) / blur)).sum(1))
    lhs_g = -blur * np.log((a[:, None] * np.exp((f0[:, None] - C) / blur)).sum(0))
    assert np.abs(lhs_f - f0).max() <= 1e-3 and np.abs(lhs_g - g0).max() <= 1e-3


@pytest.mark.parametrize("seed,n,m", [(3, 5, 7), (4, 8, 8)])
def test_small_blur_value_approaches_the_exact_emd(seed, n, m):
    a, b, C = _problem(seed, n, m)
    blur = 0.01
    eps = gr.epsilon_schedule(1, float(C.max() * 2), blur, 0.97)
    _, _, dual = gr.sinkhorn_np(a[None], b[None], C[None], eps)
    emd = _emd(a, b, C)
    # OT_eps = min <P,C> + eps KL(P | a x b) >= EMD, and the KL term of the optimal plan is at most log(1 / min a_i b_j)
    assert emd - 1e-6 <= dual[0] <= emd + blur * np.log(1.0 / (a.min() * b.min()))
