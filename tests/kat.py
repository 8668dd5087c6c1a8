"""The reference's only numeric known-answer fixture, as an op list.

Cycle text: /root/reference notebooks/tutorial.ipynb:3422-3470 (printed ExaSlang of the best individual);
expected result: notebooks/tutorial.ipynb:3373
  "solving time: 367.47..., convergence factor: 0.9212764170210773, number of iterations: 100.0".
Problem: Poisson 2D, levels 9 -> 5 (example_problems/Poisson/2D_FD_Poisson_fromL2.*)."""
from evostencils_b200 import oplist as ol

EXPECTED_CF = 0.9212764170210773
EXPECTED_ITERS = 100.0
Z = (0, 0)


def _block(bx, by):
    # key order of obtain_sympy_expression_for_local_system: first index outer (tutorial.ipynb:3433-3438)
    return tuple((0, (i, j)) for i in range(bx) for j in range(by))


def tutorial_ops():
    R = lambda l: [ol.Op(ol.OP_RESIDUAL, l, dst=ol.BUF_RES), ol.Op(ol.OP_RESTRICT, l, dst=ol.BUF_RHS, src=ol.BUF_RES)]
    ops = []
    ops += R(9) + [ol.Op(ol.OP_ZERO, 8)]                      # :3424-3426
    ops += R(8) + [ol.Op(ol.OP_ZERO, 7)]                      # :3427-3429
    ops += R(7) + [ol.Op(ol.OP_ZERO, 6)]                      # :3430-3432
    ops += [ol.Op(ol.OP_SMOOTH, 6, mode=ol.MODE_JACOBI, omega=0.35, unknowns=_block(3, 2)),   # :3433-3440
            ol.Op(ol.OP_SMOOTH, 6, mode=ol.MODE_JACOBI, omega=0.1, unknowns=_block(1, 6))]    # :3441-3448
    ops += R(6)                                               # :3449-3450
    ops += [ol.Op(ol.OP_COARSE_SOLVE, 5, count=1000, tol=1e-12)]   # :3451-3454 (CG: Poisson/2D...exa3:12-14)
    ops += [ol.Op(ol.OP_PROLONG_ADD, 6, omega=0.1), ol.Op(ol.OP_PROLONG_ADD, 7, omega=1.65),
            ol.Op(ol.OP_PROLONG_ADD, 8, omega=1.75)]          # :3455-3457
    ops += [ol.Op(ol.OP_SMOOTH, 8, mode=ol.MODE_JACOBI, omega=0.7499999999999999, unknowns=((0, Z),))]   # :3458-3460
    ops += [ol.Op(ol.OP_PROLONG_ADD, 9, omega=0.9999999999999999)]                                      # :3461
    ops += [ol.Op(ol.OP_SMOOTH, 9, mode=ol.MODE_JACOBI, omega=1.65, unknowns=((0, Z),))]                # :3462-3464
    ops += [ol.Op(ol.OP_SMOOTH, 9, mode=ol.MODE_REDBLACK, omega=1.5, unknowns=((0, Z),))]               # :3465-3470
    return ops


def drop_jacobi(ops):
    """jacobi_compat = 'exastencils_v1_1_noop' (SURVEY.md 0.5): `with jacobi` statements do nothing."""
    return [o for o in ops if not (o.code == ol.OP_SMOOTH and o.mode == ol.MODE_JACOBI)]


# Grammar string that lowers to the cycle above (reconstructed from the printed ExaSlang; weight index i
# means omega = linspace(0.1, 1.9, 37)[i], grammar/multigrid.py:428): the whole stack
# string -> tree -> lowering -> kernels can then be checked against notebooks/tutorial.ipynb:3373.
TUTORIAL_INDIVIDUAL = (
    "collective_jacobi_0(28, red_black, residual_0("
    "collective_jacobi_0(31, single, residual_0("
    "update_with_coarse_grid_correction_0(18, P_1, "
    "collective_jacobi_1(13, single, residual_1("
    "update_with_coarse_grid_correction_1(33, P_2, "
    "update_with_coarse_grid_correction_2(31, P_3, "
    "correct_with_coarse_grid_solver_3(0, P_4, CGS_4, R_3, residual_3("
    "collective_block_jacobi_3(0, ((1, 6),), residual_3("
    "collective_block_jacobi_3(5, ((3, 2),), "
    "coarsening_2(A_3, zero_3, R_2, coarsening_1(A_2, zero_2, R_1, coarsening_0(A_1, zero_1, R_0, "
    "residual_0(u_and_f))))))))))))))))))"
)
