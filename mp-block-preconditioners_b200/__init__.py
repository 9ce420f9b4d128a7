"""B200-native block preconditioner for the two-phase MAC-grid Stokes system.

The directory name follows the project (`mp-block-preconditioners_b200`); import it through the
`mp_block_preconditioners_b200` alias package at the repo root.  Modules mirror the reference's flat
layout: preconditioner.py, solve.py, apply.py, utils.py.
"""
from ._cabi import MpbpError, SIDE_LEFT, SIDE_RIGHT  # noqa: F401
from .preconditioner import (ApproxSchurOperator, MultiphaseBlockPreconditioner, Plan, SubSolver,  # noqa: F401
                             SystemOperator, thn, ths)
from .solve import (Jacobi, fgmres, gmres, last_hessenberg, main, print_true_res_norm,  # noqa: F401
                    solve_with_approx_schur_pc, solve_with_exact_schur_pc, solve_without_pc, spectral_diagnostics)
from .utils import (check_individual_operators, fill_sol_and_RHS_vecs, max_norm, print_norms,  # noqa: F401
                    weighted_L1, weighted_L2)
