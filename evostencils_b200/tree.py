"""A minimal, independent producer of cycle trees for places where neither the reference package
nor DEAP is installed (the GPU box, the test-suite, ``bench.py``).

It provides
* lightweight node classes carrying exactly the attributes the lowering reads; their class *names*
  match the reference's ``evostencils.ir`` classes (the lowering is duck-typed by name), their
  implementation is unrelated to it;
* :func:`grammar_context`: the production functions of the multigrid grammar under the names the
  reference's primitive set gives them (reference: evostencils/grammar/multigrid.py:238-385,
  SURVEY.md Appendix F), so that a saved individual string such as
  ``hof_*/individual_*.txt`` (scripts/optimize.py:175-179) can be turned into a tree with
  ``eval(string, context)`` exactly like
  ``Optimizer.generate_and_evaluate_program_from_grammar_representation`` does (program.py:919-922);
* :func:`random_individual`: a seeded grow-style generator of grammar-valid strings (what
  ``genGrow``, grammar/gp.py:6-52, does on the DEAP primitive set) for population benchmarks.

When the reference package *is* importable, trees built by it are lowered directly; a CPU test
checks that both producers lower to the same op list for the same strings.
"""
from __future__ import annotations

import itertools
import random
from dataclasses import dataclass, field
from functools import reduce
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import oplist as ol
from .problems import Problem


# ------------------------------------------------------------------------------------------------
# nodes (names mirror evostencils.ir; see module docstring)
@dataclass(frozen=True)
class Grid:
    size: Tuple[int, ...]
    spacing: Tuple[float, ...]
    level: int


class _Stencil:
    def __init__(self, entries):
        self.entries = tuple(entries)


class _PeriodicStencil:
    def __init__(self, constant_stencils, dimension):
        self.constant_stencils = constant_stencils
        self.dimension = dimension


class Single:
    @staticmethod
    def get_name():
        return "single"


class RedBlack:
    @staticmethod
    def get_name():
        return "red_black"


class ScalarOperator:
    """One (i, j) block of a system operator with a constant stencil."""

    def __init__(self, name, grid, entries):
        self.name, self.grid, self._entries = name, grid, list(entries)

    def generate_stencil(self):
        return _Stencil(self._entries)


class ZeroOperator(ScalarOperator):
    def __init__(self, grid):
        super().__init__("0", grid, [])


class InterGridEntry(ScalarOperator):
    def __init__(self, name, fine_grid, coarse_grid, entries):
        super().__init__(name, fine_grid, entries)
        self.fine_grid, self.coarse_grid = fine_grid, coarse_grid


class Restriction(InterGridEntry):
    pass


class Prolongation(InterGridEntry):
    pass


class Operator:
    """System operator: ``entries[i][j]``."""

    def __init__(self, name, entries):
        self.name, self.entries = name, entries

    @property
    def grid(self):
        return [e.grid for e in self.entries[0]]


class InterGridOperator(Operator):
    pass


class SystemRestriction(InterGridOperator):
    def __init__(self, name, entries):
        super().__init__(name, entries)

    @property
    def grid(self):
        return [row[i].coarse_grid for i, row in enumerate(self.entries)]


SystemRestriction.__name__ = "Restriction"


class SystemProlongation(InterGridOperator):
    @property
    def grid(self):
        return [row[i].fine_grid for i, row in enumerate(self.entries)]


SystemProlongation.__name__ = "Prolongation"


class Approximation:
    def __init__(self, name, grids):
        self.name, self._grids = name, list(grids)
        self.entries = [None] * len(self._grids)

    @property
    def grid(self):
        return self._grids

    @property
    def predecessor(self):
        return None


class RightHandSide(Approximation):
    pass


class ZeroApproximation(Approximation):
    def __init__(self, grids):
        super().__init__("0", grids)


class _Unary:
    def __init__(self, operand):
        self.operand = operand

    @property
    def grid(self):
        return self.operand.grid


class Diagonal(_Unary):               # decoupled Jacobi (reference: ir/smoother.py:5-6)
    pass


class ElementwiseDiagonal(_Unary):    # collective Jacobi (ir/smoother.py:9-10)
    pass


class Inverse(_Unary):
    pass


class CoarseGridSolver:
    def __init__(self, operator):
        self.name, self.operator = "Coarse-Grid Solver", operator

    @property
    def grid(self):
        return self.operator.grid


class Residual:
    def __init__(self, operator, approximation, rhs):
        self.operator, self.approximation, self.rhs = operator, approximation, rhs

    @property
    def grid(self):
        return self.rhs.grid


class Multiplication:
    def __init__(self, operand1, operand2):
        self.operand1, self.operand2 = operand1, operand2

    @property
    def grid(self):
        return self.operand1.grid


class Addition:
    def __init__(self, operand1, operand2):
        self.operand1, self.operand2 = operand1, operand2

    @property
    def grid(self):
        return self.operand1.grid


class Subtraction(Addition):
    pass


class Jacobian(_Unary):       # Newton linearisation marker of the FAS smoother (reference: ir/system.py:132-138)
    def __init__(self, operand, n_newton_steps):
        super().__init__(operand)
        self.n_newton_steps = n_newton_steps


class Cycle:
    def __init__(self, approximation, rhs, correction=None, partitioning=Single, relaxation_factor=1.0,
                 predecessor=None):
        self.approximation, self.rhs, self.correction = approximation, rhs, correction
        self.partitioning, self.relaxation_factor, self.predecessor = partitioning, relaxation_factor, predecessor

    @property
    def grid(self):
        return self.approximation.grid


# ------------------------------------------------------------------------------------------------
def _grids(problem: Problem, level: int) -> List[Grid]:
    n = 1 << level
    return [Grid((n,) * problem.dim, (1.0 / n,) * problem.dim, level) for _ in range(problem.n_fields)]


def _problem_cache(problem: Problem, key, factory):
    """Immutable terminals (operators, inter-grid operators, block splittings) are built once per problem object and
    shared by every context / individual -- like the reference, whose terminals live in the primitive set of a run
    (grammar/multigrid.py:74-137).  Only the nodes the productions MUTATE (approximations, cycles) are fresh per context."""
    try:
        cache = problem.__dict__.setdefault("_tree_terminals", {})
        # run-time parameters that enter the operators (wave number of the k / 2k / 4k runs, user parameters)
        params = getattr(problem, "parameters", None)
        key = key + (getattr(problem, "wave_number", None), tuple(sorted(params.items())) if isinstance(params, dict) else None)
    except (AttributeError, TypeError):
        return factory()
    hit = cache.get(key)
    if hit is None:
        hit = cache[key] = factory()
    return hit


def system_operator(problem: Problem, level: int, name: str) -> Operator:
    return _problem_cache(problem, ("A", level, name), lambda: _system_operator(problem, level, name))


def _system_operator(problem: Problem, level: int, name: str) -> Operator:
    table = problem.operator(level)
    grids = _grids(problem, level)
    rows = []
    for i in range(problem.n_fields):
        row = []
        for j in range(problem.n_fields):
            ent = [(ol.stencil_offset(p, problem.dim), table[i, j, p]) for p in range(ol.STENCIL_POINTS)
                   if table[i, j, p] != 0]
            row.append(ScalarOperator(f"{name}_{i}{j}", grids[j], ent) if ent else ZeroOperator(grids[j]))
        rows.append(row)
    return Operator(name, rows)


def _transfer(problem: Problem, level: int, kind: str, name: str):
    return _problem_cache(problem, ("T", level, kind, name), lambda: _make_transfer(problem, level, kind, name))


def _make_transfer(problem: Problem, level: int, kind: str, name: str):
    fine, coarse = _grids(problem, level), _grids(problem, level - 1)
    w = problem.restrict_weights() if kind == "R" else problem.prolong_weights()
    ent = [(ol.stencil_offset(p, problem.dim), w[p]) for p in range(ol.STENCIL_POINTS) if w[p] != 0]
    cls, syscls = (Restriction, SystemRestriction) if kind == "R" else (Prolongation, SystemProlongation)
    nf = problem.n_fields
    rows = []
    for i in range(nf):
        fld = problem.fields[i]
        nm = f"gen_restrictionForRes_{fld}" if kind == "R" else f"gen_prolongationForSol_{fld}"
        rows.append([cls(nm, fine[i], coarse[i], ent if i == j else []) for j in range(nf)])
    return syscls(name, rows)


def block_jacobi_operator(operator: Operator, block_shapes: Sequence[Tuple[int, ...]], dim: int) -> Operator:
    """Block-diagonal splitting with period = block shape of the row field: at periodic position k an
    entry with offset o survives iff k + o lies inside the block (what the reference's
    ``generate_collective_block_jacobi`` builds with periodic stencils, ir/smoother.py:13-22)."""
    rows = []
    for i, row in enumerate(operator.entries):
        shape = tuple(int(s) for s in block_shapes[i])
        new_row = []
        for j, entry in enumerate(row):
            base = entry.generate_stencil().entries

            def build(prefix, d):
                if d == dim:
                    kept = [(o, v) for o, v in base
                            if all(0 <= prefix[a] + o[a] < shape[a] for a in range(dim))]
                    return _Stencil(kept)
                return tuple(build(prefix + (k,), d + 1) for k in range(shape[d]))

            blk = ScalarOperator(f"{operator.name}_{i}{j}_block_diag", entry.grid, [])
            periodic = _PeriodicStencil(build((), 0), dim)
            blk.generate_stencil = (lambda p=periodic: p)
            new_row.append(blk)
        rows.append(new_row)
    return Operator(f"{operator.name}_block_diag", rows)


def block_shape_terminals(n_fields: int, dim: int, maximum_local_system_size: int):
    """All block-shape tuples the grammar offers (grammar/multigrid.py:388-407)."""
    per_field = [list(itertools.product(range(1, maximum_local_system_size + 1), repeat=dim)) for _ in range(n_fields)]
    out = []
    for perm in itertools.product(*per_field):
        terms = sum(reduce(lambda a, b: a * b, shape) for shape in perm)
        if n_fields < terms <= maximum_local_system_size:
            out.append(tuple(perm))
    return out


RELAXATION_FACTORS = np.linspace(0.1, 1.9, 37)      # grammar/multigrid.py:428


# ------------------------------------------------------------------------------------------------
def grammar_context(problem: Problem, min_level: Optional[int] = None, max_level: Optional[int] = None,
                    relaxation_factors=RELAXATION_FACTORS, fas: Optional[bool] = None) -> Dict[str, object]:
    """Name -> production function / terminal, for ``eval(individual_string, context)``.

    Fresh terminals on every call: the productions mutate and alias nodes (like the reference's
    closures, grammar/multigrid.py:257-284), so a context must not be shared between evaluations."""
    min_level = problem.min_level if min_level is None else min_level
    max_level = problem.max_level if max_level is None else max_level
    depth_total = max_level - min_level
    assert depth_total >= 1
    if fas is None:
        fas = problem.kind == ol.PROBLEM_FAS
    ctx: Dict[str, object] = {"single": Single, "red_black": RedBlack}
    approximation = Approximation("x", _grids(problem, max_level))
    rhs = RightHandSide("b", _grids(problem, max_level))
    ctx["u_and_f"] = (approximation, rhs)
    operators = {d: system_operator(problem, max_level - d, f"A_{d}") for d in range(depth_total + 1)}

    def make_level(d: int, coarsest: bool):
        level = max_level - d
        A, A_c = operators[d], operators[d + 1]
        if not coarsest:
            ctx[f"zero_{d + 1}"] = ZeroApproximation(_grids(problem, level - 1))
            ctx[f"A_{d + 1}"] = A_c
        ctx[f"P_{d + 1}"] = _transfer(problem, level, "P", f"P_{d + 1}")
        ctx[f"R_{d}"] = _transfer(problem, level, "R", f"R_{d}")

        def residual(state):
            approx, f = state
            return Cycle(approx, f, Residual(A, approx, f), predecessor=approx.predecessor)

        def update(weight_index, partitioning, cycle):
            cycle.relaxation_factor = relaxation_factors[weight_index]
            cycle.partitioning = partitioning
            return cycle, cycle.rhs

        def smoothing(weight_index, partitioning, make_splitting, cycle):
            assert isinstance(cycle.correction, Residual), "Invalid production: expected residual"
            cycle.correction = Multiplication(Inverse(make_splitting(cycle.correction.operator)), cycle.correction)
            return update(weight_index, partitioning, cycle)

        def decoupled_jacobi(weight_index, partitioning, cycle):
            return smoothing(weight_index, partitioning, Diagonal, cycle)

        def collective_jacobi(weight_index, partitioning, cycle):
            return smoothing(weight_index, partitioning, ElementwiseDiagonal, cycle)

        def collective_block_jacobi(weight_index, block_shape, cycle):
            shape_key = tuple(tuple(int(v) for v in sh) for sh in block_shape)
            return smoothing(weight_index, Single,
                             lambda op: _problem_cache(problem, ("B", id(op), shape_key),
                                                       lambda: block_jacobi_operator(op, block_shape, problem.dim)), cycle)

        def restrict(restriction, cycle):
            if fas:   # coarse rhs = R r + A_c (R u): grammar/multigrid.py:287-293
                cycle.correction = Addition(Multiplication(restriction, cycle.correction),
                                            Multiplication(A_c, Multiplication(restriction, cycle.approximation)))
            else:
                cycle.correction = Multiplication(restriction, cycle.correction)
            return cycle

        def jacobi_picard(weight_index, partitioning, cycle):
            return smoothing(weight_index, partitioning, ElementwiseDiagonal, cycle)

        def jacobi_newton(weight_index, partitioning, n_newton_steps, cycle):
            return smoothing(weight_index, partitioning,
                             lambda op: Addition(ElementwiseDiagonal(op), Jacobian(op, n_newton_steps)), cycle)

        def coarsening(coarse_operator, coarse_approximation, restriction, cycle):
            cycle = restrict(restriction, cycle)
            new_cycle = Cycle(coarse_approximation, cycle.correction,
                              Residual(coarse_operator, coarse_approximation, cycle.correction))
            new_cycle.predecessor = cycle
            return new_cycle

        def update_with_coarse_grid_correction(weight_index, prolongation, state, restriction=None):
            cycle = state[0]
            if fas:   # e_c = u_c - R u_f  (grammar/multigrid.py:275-284)
                correction = Multiplication(prolongation, Subtraction(
                    cycle, Multiplication(restriction, cycle.predecessor.approximation)))
            else:
                correction = Multiplication(prolongation, cycle)
            cycle.predecessor.correction = correction
            return update(weight_index, Single, cycle.predecessor)

        def correct_with_coarse_grid_solver(weight_index, prolongation, coarse_grid_solver, restriction, cycle):
            cycle = restrict(restriction, cycle)
            if fas:   # grammar/multigrid.py:335-340
                approximation_c = Multiplication(coarse_grid_solver, cycle.correction)
                cycle.correction = Multiplication(prolongation, Subtraction(
                    approximation_c, Multiplication(restriction, cycle.approximation)))
            else:
                cycle.correction = Multiplication(coarse_grid_solver, cycle.correction)
                cycle.correction = Multiplication(prolongation, cycle.correction)
            return update(weight_index, Single, cycle)

        ctx[f"residual_{d}"] = residual
        if problem.n_fields > 1:
            ctx[f"decoupled_jacobi_{d}"] = decoupled_jacobi
        if fas:   # grammar/multigrid.py:356-364
            ctx[f"jacobi_picard_{d}"] = jacobi_picard
            ctx[f"jacobi_newton_{d}"] = jacobi_newton
        else:
            ctx[f"collective_jacobi_{d}"] = collective_jacobi
            ctx[f"collective_block_jacobi_{d}"] = collective_block_jacobi
        if not coarsest:
            ctx[f"update_with_coarse_grid_correction_{d}"] = update_with_coarse_grid_correction
            ctx[f"coarsening_{d}"] = coarsening
        else:
            ctx[f"correct_with_coarse_grid_solver_{d}"] = correct_with_coarse_grid_solver
            ctx[f"CGS_{d + 1}"] = CoarseGridSolver(A_c)

    for d in range(depth_total):
        make_level(d, coarsest=(d == depth_total - 1))
    return ctx


def build_tree(problem: Problem, individual: str, min_level: Optional[int] = None, max_level: Optional[int] = None):
    """``expression, rhs = eval(individual, context)`` on a fresh context."""
    ctx = grammar_context(problem, min_level, max_level)
    expression, rhs = eval(individual, {"__builtins__": {}}, ctx)   # noqa: S307 - grammar strings only
    return expression


# ------------------------------------------------------------------------------------------------
def _v_cycle_tail(levels, pre, post, w, smoother, part, cw, d, presmoothed, fresh):
    """Continue building the V-cycle string from level d downward, then back up to level 0."""
    def sm(dd, arg_is_c, arg):
        return f"{smoother}_{dd}({w}, {part}, {arg if arg_is_c else f'residual_{dd}({arg})'})"

    # state on level d: either an S_d string (presmoothed) or a C_d string (fresh, unsmoothed)
    s_state, c_state = presmoothed, fresh
    while True:
        if s_state is None:
            # no pre-smoothing: go on with the fresh correction state
            c = c_state
        else:
            c = f"residual_{d}({s_state})"
        if d < levels - 1:
            c_next = f"coarsening_{d}(A_{d + 1}, zero_{d + 1}, R_{d}, {c})"
            d += 1
            if pre > 0:
                s_state = sm(d, True, c_next)
                for _ in range(pre - 1):
                    s_state = sm(d, False, s_state)
                c_state = None
            else:
                s_state, c_state = None, c_next
        else:
            s_state = f"correct_with_coarse_grid_solver_{d}({cw}, P_{d + 1}, CGS_{d + 1}, R_{d}, {c})"
            break
    # upward
    while True:
        for _ in range(post):
            s_state = sm(d, False, s_state)
        if d == 0:
            return s_state
        d -= 1
        s_state = f"update_with_coarse_grid_correction_{d}({cw}, P_{d + 1}, {s_state})"


def fas_v_cycle_individual(levels: int, pre: int = 2, post: int = 2, weight_index: int = 14, newton_steps: int = 1,
                           partitioning: str = "single", cgc_weight_index: int = 18) -> str:
    """FAS V(pre, post) cycle as a grammar string (productions of grammar/multigrid.py:360-375)."""
    def sm(d, arg_is_c, arg):
        inner = arg if arg_is_c else f"residual_{d}({arg})"
        if newton_steps > 0:
            return f"jacobi_newton_{d}({weight_index}, {partitioning}, {newton_steps}, {inner})"
        return f"jacobi_picard_{d}({weight_index}, {partitioning}, {inner})"

    s_state, c_state, d = "u_and_f", None, 0
    for _ in range(pre):
        s_state = sm(0, False, s_state)
    while True:
        c = c_state if s_state is None else f"residual_{d}({s_state})"
        if d < levels - 1:
            c_next = f"coarsening_{d}(A_{d + 1}, zero_{d + 1}, R_{d}, {c})"
            d += 1
            if pre > 0:
                s_state = sm(d, True, c_next)
                for _ in range(pre - 1):
                    s_state = sm(d, False, s_state)
                c_state = None
            else:
                s_state, c_state = None, c_next
        else:
            s_state = f"correct_with_coarse_grid_solver_{d}({cgc_weight_index}, P_{d + 1}, CGS_{d + 1}, R_{d}, {c})"
            break
    while True:
        for _ in range(post):
            s_state = sm(d, False, s_state)
        if d == 0:
            return s_state
        d -= 1
        s_state = f"update_with_coarse_grid_correction_{d}({cgc_weight_index}, P_{d + 1}, {s_state}, R_{d})"


def v_cycle_individual(levels: int, pre: int = 1, post: int = 1, weight_index: int = 18,
                       smoother: str = "collective_jacobi", partitioning: str = "red_black",
                       cgc_weight_index: int = 18) -> str:
    """Grammar string of a V(pre, post) cycle over ``levels`` coarsenings."""
    s = "u_and_f"
    for _ in range(pre):
        s = f"{smoother}_0({weight_index}, {partitioning}, residual_0({s}))"
    return _v_cycle_tail(levels, pre, post, weight_index, smoother, partitioning, cgc_weight_index, 0, s, None)


# ------------------------------------------------------------------------------------------------
def random_individual(problem: Problem, rng: random.Random, min_level: Optional[int] = None,
                      max_level: Optional[int] = None, maximum_local_system_size: int = 4, size_limit: int = 150,
                      max_smoothing_steps: int = 3, block_probability: float = 0.25) -> str:
    """A random grammar-valid individual (string).  Unlike DEAP's typed grow initialisation, which the
    reference uses (grammar/gp.py:6-52), this walks the grammar directly: at every solution state it
    applies 0..max_smoothing_steps random smoothers, then either coarsens (always, until the coarsest
    level has been visited once -- the guard types of grammar/multigrid.py:350-384 enforce the same) or
    returns upward with a coarse-grid correction.  Trees larger than ``size_limit`` nodes are rejected
    like in genGrow."""
    min_level = problem.min_level if min_level is None else min_level
    max_level = problem.max_level if max_level is None else max_level
    levels = max_level - min_level
    shapes = block_shape_terminals(problem.n_fields, problem.dim, maximum_local_system_size)

    def smoother_call(d, arg, arg_is_c):
        inner = arg if arg_is_c else f"residual_{d}({arg})"
        w = rng.randrange(37)
        r = rng.random()
        if shapes and r < block_probability:
            return f"collective_block_jacobi_{d}({w}, {rng.choice(shapes)!r}, {inner})", 4
        kinds = ["collective_jacobi"] + (["decoupled_jacobi"] if problem.n_fields > 1 else [])
        part = rng.choice(["single", "red_black"])
        return f"{rng.choice(kinds)}_{d}({w}, {part}, {inner})", 4

    while True:
        nodes = 1
        d = 0
        s_state, c_state = "u_and_f", None
        ok = True
        # downward leg
        while True:
            for _ in range(rng.randint(0, max_smoothing_steps)):
                if s_state is None:
                    s_state, k = smoother_call(d, c_state, True)
                    c_state = None
                else:
                    s_state, k = smoother_call(d, s_state, False)
                nodes += k
            c = c_state if s_state is None else f"residual_{d}({s_state})"
            nodes += 1
            if d < levels - 1:
                c_state = f"coarsening_{d}(A_{d + 1}, zero_{d + 1}, R_{d}, {c})"
                s_state = None
                nodes += 4
                d += 1
            else:
                w = rng.randrange(37)
                s_state = f"correct_with_coarse_grid_solver_{d}({w}, P_{d + 1}, CGS_{d + 1}, R_{d}, {c})"
                nodes += 5
                break
        # upward leg
        while True:
            for _ in range(rng.randint(0, max_smoothing_steps)):
                s_state, k = smoother_call(d, s_state, False)
                nodes += k
            if d == 0:
                break
            d -= 1
            s_state = f"update_with_coarse_grid_correction_{d}({rng.randrange(37)}, P_{d + 1}, {s_state})"
            nodes += 3
        if ok and nodes <= size_limit:
            return s_state
