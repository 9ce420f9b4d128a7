"""Mirror of the reference's `utils.py` for the hot path: error norms (utils.py:7-26) and the
manufactured solution / right-hand side (utils.py:159-210), on numpy arrays or torch CUDA tensors."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._cabi import check

PI = np.pi


def _is_t(a):
    return isinstance(a, torch.Tensor)


def weighted_L2(a, b, w):
    """utils.py:7-9."""
    q = a - b
    if _is_t(q):
        return float(torch.sqrt((w * q * q).sum()))
    return np.sqrt((w * q * q).sum())


def weighted_L1(a, b, w):
    """utils.py:11-14."""
    q = abs(a - b)
    return float((w * q).sum()) if _is_t(q) else (w * q).sum()


def max_norm(a, b):
    """utils.py:16-17."""
    q = abs(a - b)
    return float(q.max()) if _is_t(q) else q.max()


def device_norms(plan, a, b, w):
    """(weighted_L1, weighted_L2, max_norm) of a-b in one fused device reduction (mpbp_wnorms)."""
    ta, _ = plan._to_dev(a, 5 * plan.N)
    tb, _ = plan._to_dev(b, 5 * plan.N)
    out = (C.c_double * 3)()
    with torch.cuda.device(plan.device):
        check(plan.lib.mpbp_wnorms(plan.h, ta.data_ptr(), tb.data_ptr(), ta.numel(), float(w), out, plan.stream()))
    return out[0], out[1], out[2]


def print_norms(u_approx, u_vec, dx, dy, n, show_max=True):
    """utils.py:19-26."""
    L1_norm = weighted_L1(u_approx, u_vec, dx * dy)
    L2_norm = weighted_L2(u_approx, u_vec, dx * dy)
    print(f"The L1_norm for n = {n} is {L1_norm}")
    print(f"The L2_norm for n = {n} is {L2_norm}")
    if show_max:
        Max_norm = max_norm(u_approx, u_vec)
        print(f"The max_norm for n = {n} is {Max_norm}")
    return L1_norm, L2_norm


def fill_sol_and_RHS_vecs(n, u_n_x_fcn, u_n_y_fcn, u_s_x_fcn, u_s_y_fcn, p_fcn, b_n_x_fcn, b_n_y_fcn, b_s_x_fcn,
                          b_s_y_fcn, b_p_fcn):
    """utils.py:159-210 with the per-cell Python loop replaced by one vectorised evaluation of the
    caller's functions at the same sample points (u: left faces, v: top faces, p: cell centres)."""
    h = 1 / n
    r = np.arange(n, dtype=np.float64)[:, None] + np.zeros((1, n))
    c = np.arange(n, dtype=np.float64)[None, :] + np.zeros((n, 1))
    yu, xu = -(r + 0.5) * h, c * h            # utils.py:187
    yv, xv = -r * h, (c + 0.5) * h            # utils.py:188
    yp, xp = -(r + 0.5) * h, (c + 0.5) * h    # utils.py:193

    def ev(f, y, x):
        return (np.asarray(f(y, x), dtype=np.float64) + np.zeros((n, n))).ravel()

    u_vec = np.concatenate([ev(u_n_x_fcn, yu, xu), ev(u_n_y_fcn, yv, xv), ev(u_s_x_fcn, yu, xu),
                            ev(u_s_y_fcn, yv, xv), ev(p_fcn, yp, xp)])
    b_vec = np.concatenate([ev(b_n_x_fcn, yu, xu), ev(b_n_y_fcn, yv, xv), ev(b_s_x_fcn, yu, xu),
                            ev(b_s_y_fcn, yv, xv), ev(b_p_fcn, yp, xp)])
    return u_vec, b_vec


def manufactured_device(plan, b_p_sign=-1.0):
    """(u_vec, b_vec) of solve.main (solve.py:52-81) assembled on the GPU as torch tensors."""
    u = torch.empty(5 * plan.N, dtype=torch.float64, device=plan.device)
    b = torch.empty_like(u)
    with torch.cuda.device(plan.device):
        check(plan.lib.mpbp_fill_manufactured(plan.h, u.data_ptr(), b.data_ptr(), float(b_p_sign), plan.stream()))
    return u, b


def check_individual_operators(n, xi, L_n, D_n, XI_n, G_n, check_laplacian_op=True, check_divergence_op=True,
                               check_xi_op=True, check_gradient_op=True, verbose=True):
    """utils.check_individual_operators (utils.py:42-157): truncation error of each block operator applied to the
    reference's analytic fields (thn = 1/4 sin 2 pi x sin 2 pi y + 1/2), printed the way the reference prints them.
    The operators are the objects returned by `get_block_matrices` (they run on the GPU); the per-cell Python loops are
    one vectorised evaluation.  Returns {"D": (L1, L2), "G": ..., "XI": ..., "L": ...} for the requested checks."""
    from .preconditioner import thn, ths
    h = 1 / n
    r = np.arange(n, dtype=np.float64)[:, None] + np.zeros((1, n))
    c = np.arange(n, dtype=np.float64)[None, :] + np.zeros((n, 1))
    yu, xu, yv, xv, yp, xp = -(r + .5) * h, c * h, -r * h, (c + .5) * h, -(r + .5) * h, (c + .5) * h
    ux = lambda y, x: np.sin(2 * PI * x) * np.cos(2 * PI * y)   # utils.py:53
    uy = lambda y, x: np.cos(2 * PI * x) * np.sin(2 * PI * y)   # utils.py:54
    u_n = np.concatenate([ux(yu, xu).ravel(), uy(yv, xv).ravel()])
    p_n = ux(yp, xp).ravel()                                    # utils.py:55
    w = h * h
    out = {}

    def report(name, exact, approx):
        L2, L1 = weighted_L2(exact, approx, w), weighted_L1(exact, approx, w)
        out[name] = (L1, L2)
        if verbose:
            print(f"Printing error norms for application of {name}:")
            print(f"The L1_norm for n = {n} is {L1}")
            print(f"The L2_norm for n = {n} is {L2}\n\n")
    if check_divergence_op:
        exact = (2 * PI * np.cos(2 * PI * xp) * np.cos(2 * PI * yp) + 1 / 2 * PI * np.sin(4 * PI * xp) * np.sin(4 * PI * yp)).ravel()  # :60
        report("D", exact, D_n @ u_n)
    if check_gradient_op:
        gx = lambda y, x: PI / 2 * np.sin(2*PI*x) * np.sin(2*PI*y) * np.cos(2*PI*x) * np.cos(2*PI*y) + PI * np.cos(2*PI*x) * np.cos(2*PI*y)  # :85
        gy = lambda y, x: -PI / 2 * np.sin(2*PI*x)**2 * np.sin(2*PI*y)**2 - PI * np.sin(2*PI*x) * np.sin(2*PI*y)                              # :86
        report("G", np.concatenate([gx(yu, xu).ravel(), gy(yv, xv).ravel()]), G_n @ p_n)
    if check_xi_op:
        exact = np.concatenate([(xi * thn(yu, xu) * ths(yu, xu) * ux(yu, xu)).ravel(),
                                (xi * thn(yv, xv) * ths(yv, xv) * uy(yv, xv)).ravel()])  # :122-123
        report("XI", exact, XI_n @ u_n)
    if check_laplacian_op:
        lx = lambda y, x: -4*PI*PI*np.sin(2*PI*x)**2*np.sin(2*PI*y)*np.cos(2*PI*y) - 4*PI*PI*np.sin(2*PI*x)*np.cos(2*PI*y)  # :135
        ly = lambda y, x: -4*PI*PI*np.sin(2*PI*x)*np.cos(2*PI*x)*np.sin(2*PI*y)**2 - 4*PI*PI*np.cos(2*PI*x)*np.sin(2*PI*y)  # :136
        report("L", np.concatenate([lx(yu, xu).ravel(), ly(yv, xv).ravel()]), L_n @ u_n)
    return out
