"""Lowering of an ``evostencils.ir`` cycle tree to an op list (:class:`evostencils_b200.oplist.Program`).

This is the counterpart of the reference's text emitter ``ProgramGenerator.generate_multigrid``
(reference: evostencils/code_generation/exastencils.py:684-925) and of ``generate_cycle_function``
(:318-336): same recursion, same emission order, but it produces statements as data instead of
ExaSlang text, and it never mutates the tree (the reference sets and resets ``.valid`` flags,
:330-334, :702; here a local set of node ids plays that role).

The tree is duck-typed by class *name* so that it accepts the reference's own ``evostencils.ir``
objects (drop-in use below ``Optimizer``) as well as the lightweight nodes of
:mod:`evostencils_b200.tree` (used where the reference package is not installed).

Local systems of smoothers (``solve locally``) are derived without sympy: the reference builds the
equations symbolically (ir/transformations.py:51-145) only to find out *which cells are solved
together*; :func:`local_system_statements` computes the same key sets and the same
dependent/independent split from the stencil offsets alone.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import oplist as ol


# ------------------------------------------------------------------------------------------------
# duck typing helpers
_NAMES: Dict[type, frozenset] = {}


def _names(obj) -> frozenset:
    t = type(obj)
    names = _NAMES.get(t)
    if names is None:
        names = _NAMES[t] = frozenset(c.__name__ for c in t.__mro__)
    return names


def _is(obj, name: str) -> bool:
    return name in _names(obj)


# Results that depend only on an operator OBJECT (its stencil tables, the statements of a smoother built on it) are
# kept per object: the terminals of a grammar are shared by all individuals of a run (the reference builds them once,
# grammar/multigrid.py:74-137; tree.py caches them per problem), so a generation lowers each of them once.  The cache
# holds a reference to the object, which keeps its id() unique.
_PER_OBJECT: Dict[Tuple, Tuple[object, object]] = {}


def _per_object(kind: str, obj, extra, compute):
    key = (kind, id(obj), extra)
    hit = _PER_OBJECT.get(key)
    if hit is not None and hit[0] is obj:
        return hit[1]
    if len(_PER_OBJECT) > 4096:        # individuals built from uncached terminals: do not grow without bound
        _PER_OBJECT.clear()
    value = compute()
    _PER_OBJECT[key] = (obj, value)
    return value


def _is_system_approximation(obj) -> bool:
    """isinstance(obj, system.Approximation) -- includes system.RightHandSide / ZeroApproximation."""
    return _is(obj, "Approximation") and hasattr(obj, "entries")


def _is_system_rhs(obj) -> bool:
    return _is(obj, "RightHandSide") and hasattr(obj, "entries")


def _is_zero_approximation(obj) -> bool:
    return _is(obj, "ZeroApproximation") and hasattr(obj, "entries")


def _is_intergrid(obj) -> bool:
    return _is(obj, "InterGridOperator") and hasattr(obj, "entries")


def _level(grid) -> int:
    return int(grid.level)


# ------------------------------------------------------------------------------------------------
# stencils
def _constant_entries(stencil) -> List[Tuple[Tuple[int, ...], complex]]:
    """Entries of a constant stencil; a periodic stencil of period 1 is unwrapped."""
    if stencil is None:
        return []
    while hasattr(stencil, "constant_stencils"):
        inner = stencil.constant_stencils
        while isinstance(inner, tuple):
            if len(inner) != 1:
                raise ValueError("periodic stencil where a constant stencil is expected")
            inner = inner[0]
        stencil = inner
        if stencil is None:
            return []
    out = []
    for offset, value in stencil.entries:
        out.append((tuple(int(o) for o in offset), complex(value)))
    return out


def operator_table(system_operator, n_fields: int) -> np.ndarray:
    """Coefficient table [nf, nf, 27] of a system operator (``system.Operator`` of stencil expressions,
    built by grammar/multigrid.py:74-137 from the problem's equations); read-only, cached per operator object."""
    def compute():
        table = _operator_table(system_operator, n_fields)
        table.setflags(write=False)
        return table
    return _per_object("A", system_operator, n_fields, compute)


def _operator_table(system_operator, n_fields: int) -> np.ndarray:
    table = np.zeros((n_fields, n_fields, ol.STENCIL_POINTS), dtype=np.complex128)
    for i, row in enumerate(system_operator.entries):
        for j, entry in enumerate(row):
            if _is(entry, "ZeroOperator"):
                continue
            for offset, value in _constant_entries(entry.generate_stencil()):
                table[i, j, ol.stencil_index(offset)] += value
    if np.all(table.imag == 0.0):
        return table.real.copy()
    return table


def transfer_table(intergrid_system_operator) -> np.ndarray:
    """Weights (3^d table) of the restriction / prolongation of field 0 (all fields share it); read-only, cached."""
    def compute():
        w = _transfer_table(intergrid_system_operator)
        w.setflags(write=False)
        return w
    return _per_object("T", intergrid_system_operator, None, compute)


def _transfer_table(intergrid_system_operator) -> np.ndarray:
    entry = intergrid_system_operator.entries[0][0]
    w = np.zeros(ol.STENCIL_POINTS)
    for offset, value in _constant_entries(entry.generate_stencil()):
        w[ol.stencil_index(offset)] += float(np.real(value))
    return w


# ------------------------------------------------------------------------------------------------
# local systems
def _periodic_get(array, index: Tuple[int, ...]):
    """Element of a periodic (nested tuple) stencil at ``index`` with the reference's modulo rule
    (ir/transformations.py:66-89: ``array[k % len(array)]`` per dimension)."""
    cur = array
    for k in index:
        cur = cur[k % len(cur)]
    return cur


def _period(array, dim: int) -> Tuple[int, ...]:
    shape = []
    cur = array
    for _ in range(dim):
        shape.append(len(cur))
        cur = cur[0]
    return tuple(shape)


def _smoother_blocks(smoothing_operator, system_operator, n_fields: int, dim: int):
    """For every (row field i, column field j): the periodic array of stencils of the smoother's
    splitting matrix (stencil1 of ir/transformations.py:90-118) and its period."""
    zero = (0,) * dim
    blocks = [[None] * n_fields for _ in range(n_fields)]

    def constant_as_periodic(entries):
        cur = entries
        for _ in range(dim):
            cur = (cur,)
        return cur

    if type(smoothing_operator).__name__ == "Diagonal":            # system.Diagonal: decoupled Jacobi
        for i, row in enumerate(smoothing_operator.operand.entries):
            for j, entry in enumerate(row):
                if i == j:
                    diag = [(o, v) for o, v in _constant_entries(entry.generate_stencil()) if o == zero]
                else:
                    diag = []
                blocks[i][j] = constant_as_periodic(diag)
        kind = "decoupled"
    elif type(smoothing_operator).__name__ == "ElementwiseDiagonal":   # collective Jacobi
        for i, row in enumerate(smoothing_operator.operand.entries):
            for j, entry in enumerate(row):
                diag = [(o, v) for o, v in _constant_entries(entry.generate_stencil()) if o == zero]
                blocks[i][j] = constant_as_periodic(diag)
        kind = "collective"
    elif hasattr(smoothing_operator, "entries"):
        # custom splitting: block-diagonal periodic stencils (ir/smoother.py:13-38, stencils/multiple.py:204-217)
        for i, row in enumerate(smoothing_operator.entries):
            for j, entry in enumerate(row):
                st = entry.generate_stencil()
                if st is None:
                    blocks[i][j] = constant_as_periodic([])
                    continue
                if hasattr(st, "constant_stencils"):
                    def conv(node, d):
                        if d == 0:
                            return [] if node is None else [(tuple(int(o) for o in off), complex(val))
                                                            for off, val in node.entries]
                        return tuple(conv(x, d - 1) for x in node)
                    blocks[i][j] = conv(st.constant_stencils, dim)
                else:
                    blocks[i][j] = constant_as_periodic(_constant_entries(st))
        kind = "block"
    else:
        raise RuntimeError("Can not extract equations from smoothing operator")
    return blocks, kind


def local_system_statements(smoothing_operator, system_operator, n_fields: int, dim: int):
    """Cached per smoother: Diagonal / ElementwiseDiagonal are thin wrappers created per individual around the shared
    system operator (key: wrapper type + operand); block splittings are shared objects themselves (tree.py)."""
    tname = type(smoothing_operator).__name__
    if tname in ("Diagonal", "ElementwiseDiagonal"):
        anchor, extra = smoothing_operator.operand, (tname, n_fields, dim)
    else:
        anchor, extra = smoothing_operator, (tname, n_fields, dim)
    return _per_object("S", anchor, extra,
                       lambda: _local_system_statements(smoothing_operator, system_operator, n_fields, dim))


def _local_system_statements(smoothing_operator, system_operator, n_fields: int, dim: int):
    """Statements ``[(unknowns, ...)]`` a smoother expands to, in emission order.

    Mirrors exastencils.py:769-822: the keys of ``obtain_sympy_expression_for_local_system`` are
    (field, block index); an equation is *independent* iff none of the new-value symbols occurring in
    it occurs in another equation (ir/transformations.py:124-145); every independent equation becomes
    its own statement (emitted first, in key order), all dependent ones one joint statement; a
    collective (ElementwiseDiagonal) smoother is always one joint statement (:775-777)."""
    blocks, kind = _smoother_blocks(smoothing_operator, system_operator, n_fields, dim)
    keys: List[Tuple[int, Tuple[int, ...]]] = []
    new_symbols: Dict[Tuple[int, Tuple[int, ...]], set] = {}
    for i in range(n_fields):
        for j in range(n_fields):
            arr = blocks[i][j]
            period = _period(arr, dim)
            # max_period = max(len(array1), len(array2)); array2 (the system stencil) has period 1
            for index in np.ndindex(*period):
                index = tuple(int(v) for v in index)
                key = (i, index)
                if key not in new_symbols:
                    new_symbols[key] = set()
                    keys.append(key)
                for offset, value in _periodic_get(arr, index):
                    if value == 0:
                        continue
                    new_symbols[key].add((j, tuple(a + b for a, b in zip(index, offset))))
    independent, dependent = [], []
    for key in keys:
        mine = new_symbols[key]
        is_independent = True
        for other in keys:
            if other == key:
                continue
            if mine & new_symbols[other]:
                is_independent = False
                break
        (independent if is_independent else dependent).append(key)
    if kind == "collective":
        dependent = dependent + independent   # dependent_equations.extend(independent_equations) (:776)
        independent = []
    statements = [((k,),) for k in independent]
    statements = [tuple(s[0]) for s in statements]
    if dependent:
        statements.append(tuple(dependent))
    return statements


# ------------------------------------------------------------------------------------------------
class LoweringError(RuntimeError):
    pass


class Lowering:
    """One lowering run = one call of the reference's ``generate_cycle_function``."""

    def __init__(self, min_level: int, max_level: int, n_fields: int, dim: int, use_jacobi_prefix: bool = True,
                 cgs_max_iters: int = 1000, cgs_tol: float = 1e-12, cycle_registry: Optional[Dict[int, list]] = None):
        self.min_level = min_level
        self.max_level = max_level          # absolute finest level (5th argument of generate_cycle_function)
        self.nf = n_fields
        self.dim = dim
        self.use_jacobi_prefix = use_jacobi_prefix
        self.cgs_max_iters = cgs_max_iters
        self.cgs_tol = cgs_tol
        self.cycle_registry = cycle_registry or {}
        self.ops: List[ol.Op] = []
        self.operators: Dict[int, np.ndarray] = {}
        self.restrict_w: Optional[np.ndarray] = None
        self.prolong_w: Optional[np.ndarray] = None
        self._valid = set()

    # -- field selection (exastencils.py:295-316, :594-604) ------------------------------------------
    def _sol(self, level: int) -> int:
        return ol.BUF_SOL

    def _cor(self, level: int) -> int:
        # gen_error_<f> is the solution field below the finest level
        return ol.BUF_COR if level == self.max_level else ol.BUF_SOL

    def _source_buffer(self, expression, level: int) -> int:
        # NB: system.RightHandSide derives from system.Approximation, so the reference's RightHandSide
        # branch (:601-602) is dead code and a right-hand side maps to the solution field; mirrored.
        if _is_system_approximation(expression) or _is(expression, "Cycle"):
            return self._sol(level)
        if _is(expression, "Residual"):
            return ol.BUF_RES
        return self._cor(level)

    def _register_operator(self, system_operator, level: int):
        seen = self.__dict__.setdefault("_registered_operators", {})
        if seen.get(id(system_operator)) == level:        # the same terminal again (kept alive by the tree)
            return
        seen[id(system_operator)] = level
        table = operator_table(system_operator, self.nf)
        if level in self.operators:
            if not np.array_equal(self.operators[level], table):
                raise LoweringError(f"two different system operators on level {level}")
        else:
            self.operators[level] = table

    def _register_transfer(self, op):
        w = transfer_table(op)
        if _is(op, "Restriction") or _is(op.entries[0][0], "Restriction"):
            if self.restrict_w is not None and not np.array_equal(self.restrict_w, w):
                raise LoweringError("level-dependent restriction weights are not supported")
            self.restrict_w = w
        else:
            if self.prolong_w is not None and not np.array_equal(self.prolong_w, w):
                raise LoweringError("level-dependent prolongation weights are not supported")
            self.prolong_w = w

    @staticmethod
    def _grid_level(expression) -> int:
        grid = expression.grid
        if isinstance(grid, (list, tuple)):
            return _level(grid[0])
        return _level(grid)

    # -- recursion (exastencils.py:684-925) ---------------------------------------------------------------
    def emit(self, expression):
        if _is(expression, "Cycle"):
            self._emit_cycle(expression)
        elif _is(expression, "Residual"):
            self._emit_residual(expression)
        elif _is(expression, "Multiplication"):
            self._emit_multiplication(expression)
        elif _is(expression, "Addition") or _is(expression, "Subtraction"):
            # only string-joins the two sides in the reference (:914-922); meaningful for FAS only
            self.emit(expression.operand1)
            self.emit(expression.operand2)
        else:
            raise LoweringError("Not implemented")

    def _emit_cycle(self, expression):
        weight = float(expression.relaxation_factor)
        correction = expression.correction
        level = self._grid_level(expression)
        if _is(correction, "Residual"):
            # Richardson step (:698-726)
            if not _is_system_rhs(correction.rhs) and id(expression.rhs) not in self._valid:
                self.emit(expression.rhs)
                self._valid.add(id(expression.rhs))
            if not _is_system_approximation(correction.approximation):
                self.emit(expression.approximation)
            if _is_zero_approximation(expression.approximation):
                self.ops.append(ol.Op(ol.OP_ZERO, level, dst=self._sol(level)))
            self._register_operator(correction.operator, level)
            self.ops.append(ol.Op(ol.OP_RICHARDSON, level, omega=weight))
        elif _is(correction, "Multiplication"):
            op1 = correction.operand1
            if _is_intergrid(op1):
                # coarse-grid correction (:727-743)
                self.emit(correction.operand2)
                entry = op1.entries[0][0]
                if _is(entry, "Prolongation"):
                    op_level = _level(entry.coarse_grid)
                    self._register_transfer(op1)
                    src = self._source_buffer(correction.operand2, op_level)
                    if level != op_level + 1:
                        raise LoweringError("prolongation across more than one level")
                    self.ops.append(ol.Op(ol.OP_PROLONG_ADD, level, dst=self._sol(level), src=src, omega=weight))
                elif _is(entry, "Restriction"):
                    raise LoweringError("restriction as a correction is not supported")
                else:
                    raise LoweringError("Unexpected entry")
            elif _is(op1, "Inverse") or _is(op1, "KrylovSubspaceMethod"):
                residual = correction.operand2
                if not _is_system_rhs(residual.rhs) and id(residual.rhs) not in self._valid:
                    self.emit(residual.rhs)
                    self._valid.add(id(residual.rhs))
                if not _is_system_approximation(residual.approximation):
                    self.emit(residual.approximation)
                if _is_zero_approximation(expression.approximation):
                    self.ops.append(ol.Op(ol.OP_ZERO, level, dst=self._sol(level)))
                if _is(op1, "KrylovSubspaceMethod"):
                    # unreachable from generate_primitive_set (ir/krylov_subspace.py:10 cannot even be constructed)
                    raise LoweringError("Krylov subspace smoothers are not generated by the grammar")
                self._register_operator(residual.operator, level)
                self._emit_smoother(expression, op1.operand, residual.operator, level, weight)
            else:
                raise LoweringError("Unsupported operator")
        else:
            raise LoweringError("Expected multiplication")

    def _emit_smoother(self, cycle, smoothing_operator, system_operator, level, weight):
        partitioning = cycle.partitioning
        pname = partitioning.__name__ if isinstance(partitioning, type) else type(partitioning).__name__
        if pname == "Single":
            mode = ol.MODE_JACOBI if self.use_jacobi_prefix else ol.MODE_LEX
        elif pname == "RedBlack":
            mode = ol.MODE_REDBLACK
        else:
            raise LoweringError(f"partitioning {pname} does not exist in evostencils.ir.partitioning")
        for unknowns in local_system_statements(smoothing_operator, system_operator, self.nf, self.dim):
            if len(unknowns) > ol.MAX_UNKNOWNS:
                raise LoweringError("local system larger than 8 unknowns")
            self.ops.append(ol.Op(ol.OP_SMOOTH, level, mode=mode, omega=weight,
                                  unknowns=tuple((f, tuple(idx)) for f, idx in unknowns)))

    def _emit_residual(self, expression):
        level = self._grid_level(expression)
        if not _is_system_rhs(expression.rhs) and id(expression.rhs) not in self._valid:
            self.emit(expression.rhs)
            self._valid.add(id(expression.rhs))
        if not _is_system_approximation(expression.approximation):
            self.emit(expression.approximation)
        if _is_zero_approximation(expression.approximation):
            self.ops.append(ol.Op(ol.OP_ZERO, level, dst=self._sol(level)))
        self._register_operator(expression.operator, level)
        self.ops.append(ol.Op(ol.OP_RESIDUAL, level, dst=ol.BUF_RES))

    def _emit_multiplication(self, expression):
        op1 = expression.operand1
        if _is_intergrid(op1):
            # R * x -> RHS@coarse ; P * x -> COR@fine   (:855-873)
            self.emit(expression.operand2)
            entry = op1.entries[0][0]
            self._register_transfer(op1)
            out_level = self._grid_level(expression)
            if _is(entry, "Prolongation"):
                op_level = _level(entry.coarse_grid)
                src = self._source_buffer(expression.operand2, op_level)
                self.ops.append(ol.Op(ol.OP_PROLONG_SET, op_level + 1, dst=self._cor(op_level + 1), src=src))
            elif _is(entry, "Restriction"):
                op_level = _level(entry.fine_grid)
                src = self._source_buffer(expression.operand2, op_level)
                self.ops.append(ol.Op(ol.OP_RESTRICT, op_level, dst=ol.BUF_RHS, src=src))
            else:
                raise LoweringError("Unexpected entry")
            del out_level
        elif _is(op1, "CoarseGridSolver"):
            # (:874-911)
            self.emit(expression.operand2)
            level = self._grid_level(expression.operand2)
            self._register_operator(op1.operator, level)
            if level == self.min_level:
                # gen_rhs = RHS; gen_error = 0; gen_mgCycle@min()  -> Krylov solve from a zero guess
                self.ops.append(ol.Op(ol.OP_COARSE_SOLVE, level, count=self.cgs_max_iters, tol=self.cgs_tol))
            else:
                # SOL = 0; call of an already defined gen_mgCycle@level (multi-run mode): inline it
                self.ops.append(ol.Op(ol.OP_ZERO, level, dst=self._sol(level)))
                if level not in self.cycle_registry:
                    raise LoweringError(f"coarse-grid solver on level {level}: no cycle registered for that level")
                registered = self.cycle_registry[level]
                self.ops.extend(registered["ops"])
                for l, t in registered["operators"].items():
                    self.operators.setdefault(l, t)
            # COR = SOL (a no-op below the finest level where both are gen_error)
            if self._cor(level) != self._sol(level):
                self.ops.append(ol.Op(ol.OP_COPY, level, dst=self._cor(level), src=self._sol(level)))
        else:
            raise LoweringError("Not implemented")


def lower_cycle(expression, min_level: int, max_level: int, n_fields: int, dim: int, use_jacobi_prefix: bool = True,
                cgs_max_iters: int = 1000, cgs_tol: float = 1e-12, cycle_registry=None,
                default_restrict=None, default_prolong=None) -> ol.Program:
    """IR tree -> :class:`Program` (what ``generate_cycle_function`` + code generation produce)."""
    lo = Lowering(min_level, max_level, n_fields, dim, use_jacobi_prefix, cgs_max_iters, cgs_tol, cycle_registry)
    lo.emit(expression)
    prog = ol.Program(dim=dim, n_fields=n_fields, min_level=min_level, max_level=max_level, ops=lo.ops,
                      operators=lo.operators)
    prog.restrict_w = lo.restrict_w if lo.restrict_w is not None else default_restrict
    prog.prolong_w = lo.prolong_w if lo.prolong_w is not None else default_prolong
    return prog


def apply_jacobi_compat(program: ol.Program, mode: str) -> ol.Program:
    """``jacobi_compat``: 'intended' keeps `with jacobi` statements; 'exastencils_v1_1_noop' drops them,
    which is what the reference *as shipped* computes (its `advance` patch searches for "[next]" while
    ExaStencils prints "<next>": exastencils.py:348)."""
    if mode == "intended":
        return program
    if mode != "exastencils_v1_1_noop":
        raise ValueError(mode)
    import copy
    p = copy.copy(program)
    p.ops = [o for o in program.ops if not (o.code == ol.OP_SMOOTH and o.mode == ol.MODE_JACOBI)]
    return p


def optimise(program: ol.Program) -> ol.Program:
    """Peephole fusion that does not change results: RESIDUAL immediately followed by RESTRICT of that
    residual becomes RESIDUAL_RESTRICT when the stored residual is dead (overwritten before any other
    read)."""
    import copy
    ops = list(program.ops)
    out: List[ol.Op] = []
    i = 0
    while i < len(ops):
        o = ops[i]
        if (o.code == ol.OP_RESIDUAL and i + 1 < len(ops) and ops[i + 1].code == ol.OP_RESTRICT
                and ops[i + 1].level == o.level and ops[i + 1].src == ol.BUF_RES and ops[i + 1].dst == ol.BUF_RHS):
            dead = True
            for later in ops[i + 2:]:
                reads_res = (later.src == ol.BUF_RES and later.code in (ol.OP_RESTRICT, ol.OP_COPY, ol.OP_PROLONG_ADD,
                                                                        ol.OP_PROLONG_SET)
                             and (later.level == o.level if later.code in (ol.OP_RESTRICT, ol.OP_COPY)
                                  else later.level - 1 == o.level))
                if reads_res:
                    dead = False
                    break
                if later.code == ol.OP_RESIDUAL and later.level == o.level:
                    break
            if dead:
                out.append(ol.Op(ol.OP_RESIDUAL_RESTRICT, o.level, dst=ol.BUF_RHS, src=ol.BUF_RES))
                i += 2
                continue
        out.append(o)
        i += 1
    p = copy.copy(program)
    p.ops = out
    return p
