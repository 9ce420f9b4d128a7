"""Mirror of the reference's `apply.py` (operator-apply check, apply.py:8-81) with the current
constructor/factory API: b_approx = A @ u_vec against the manufactured right-hand side."""
from __future__ import annotations

from .preconditioner import MultiphaseBlockPreconditioner
from .utils import device_norms, manufactured_device


def apply_check(n=32, xi=1.0, eta_n=1.0, eta_s=1.0, c=1.0, d=-1.0, b_p_sign=-1.0, verbose=True):
    """Returns (L1, L2, max) of A u_exact - b_exact (apply.py:71-81)."""
    bp = MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s)
    A, _, _, _, _ = bp.get_big_A_matrix(c=c, d_u=d)
    u_vec, b_vec = manufactured_device(A.plan, b_p_sign)
    b_approx = A @ u_vec                                   # apply.py:72
    L1, L2, mx = device_norms(A.plan, b_vec, b_approx, (1 / n) * (1 / n))
    if verbose:
        print("Printing error norms for application of big A:")
        print(f"The L1_norm for n = {n} is {L1}")
        print(f"The L2_norm for n = {n} is {L2}")
        print(f"The max_norm for n = {n} is {mx}")
    return L1, L2, mx


if __name__ == "__main__":
    apply_check()
