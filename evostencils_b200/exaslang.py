"""ExaSlang (layer 3) text of a lowered cycle.

``Optimizer`` concatenates the return value of ``generate_cycle_function`` into ``solver_program``
and prints it as "ExaSlang Code" at the end of a run (reference: optimization/program.py:891-898,
scripts/optimize.py:162).  The B200 backend does not need the text, but keeps producing it so that
results stay consumable by ExaStencils users; statement order and statement syntax follow the
reference emitter (code_generation/exastencils.py:684-925) -- the test-suite compares the two texts
statement by statement for the golden individuals.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import oplist as ol


def _field_name(buf: int, level: int, max_level: int, field: str, rhs_name: str) -> str:
    if buf == ol.BUF_SOL:
        return f"{field}@{level}" if level == max_level else f"gen_error_{field}@{level}"
    if buf == ol.BUF_RHS:
        return f"{rhs_name}@{level}"
    if buf == ol.BUF_RES:
        return f"gen_residual_{field}@{level}"
    if buf == ol.BUF_COR:
        return f"gen_error_{field}@{level}"
    if buf == ol.BUF_APX:
        return f"gen_approximation_{field}@{level}"
    raise ValueError(buf)


def _fmt(v) -> str:
    v = complex(v)
    if v.imag == 0:
        return repr(float(v.real))
    return repr(v)


def _row_text(table: np.ndarray, i: int, fields: Sequence[str], level: int, max_level: int, anchor, dim: int) -> str:
    """Full operator row of field i at node ``anchor`` as a sum of field accesses."""
    terms = []
    for j, fld in enumerate(fields):
        name = f"{fld}@{level}" if level == max_level else f"gen_error_{fld}@{level}"
        for p in range(ol.STENCIL_POINTS):
            c = table[i, j, p]
            if c == 0:
                continue
            off = ol.stencil_offset(p, dim)
            pos = ", ".join(str(a + o) for a, o in zip(anchor, off))
            terms.append(f"{_fmt(c)}*{name}@[{pos}]")
    return " + ".join(terms).replace("+ -", "- ")


def program_to_exaslang(program: ol.Program, fields: Sequence[str], rhs_names: Sequence[str], level: int,
                        cycle_name: str = "gen_mgCycle") -> str:
    mx = program.max_level
    dim = program.dim
    out: List[str] = [f"Function {cycle_name}@{level} {{\n"]
    for op in program.ops:
        l = op.level
        if op.code == ol.OP_ZERO:
            for f, r in zip(fields, rhs_names):
                out.append(f"\t{_field_name(op.dst, l, mx, f, r)} = 0\n")
        elif op.code == ol.OP_COPY:
            for f, r in zip(fields, rhs_names):
                out.append(f"\t{_field_name(op.dst, l, mx, f, r)} = {_field_name(op.src, l, mx, f, r)}\n")
        elif op.code == ol.OP_RESIDUAL:
            for i, (f, r) in enumerate(zip(fields, rhs_names)):
                line = f"\t{_field_name(ol.BUF_RES, l, mx, f, r)} = {_field_name(ol.BUF_RHS, l, mx, f, r)}"
                for j, fj in enumerate(fields):
                    if np.any(program.operators[l][i, j] != 0):
                        line += f" - (A{i}{j}@{l}*{_field_name(ol.BUF_SOL, l, mx, fj, rhs_names[j])})"
                out.append(line + "\n")
        elif op.code == ol.OP_RICHARDSON:
            for i, (f, r) in enumerate(zip(fields, rhs_names)):
                line = f"\t{_field_name(ol.BUF_SOL, l, mx, f, r)} += {op.omega} * ({_field_name(ol.BUF_RHS, l, mx, f, r)}"
                for j, fj in enumerate(fields):
                    if np.any(program.operators[l][i, j] != 0):
                        line += f" - (A{i}{j}@{l}*{_field_name(ol.BUF_SOL, l, mx, fj, rhs_names[j])})"
                out.append(line + ")\n")
        elif op.code == ol.OP_SMOOTH:
            colored = op.mode == ol.MODE_REDBLACK
            ind = "\t" if colored else ""
            if colored:
                out.append(f"{ind}color with {{\n")
                out.append(f"{ind}\t((" + " + ".join(f"i{d}" for d in range(dim)) + ") % 2),\n")
            jac = "with jacobi " if op.mode == ol.MODE_JACOBI else ""
            f0 = fields[op.unknowns[0][0]]
            at = f"{f0}@{l}" if l == mx else f"gen_error_{f0}@{l}"
            for _ in range(max(1, op.count)):
                out.append(f"\t{ind}solve locally at {at} {jac}relax {op.omega} {{\n")
                for fi, off in op.unknowns:
                    f = fields[fi]
                    unk = (f"{f}@{l}" if l == mx else f"gen_error_{f}@{l}") + "@[" + ", ".join(str(o) for o in off) + "]"
                    rhs = f"{rhs_names[fi]}@{l}@[" + ", ".join(str(o) for o in off) + "]"
                    row = _row_text(program.operators[l], fi, fields, l, mx, off, dim)
                    out.append(f"\t\t{ind}{unk} => ({row}) == {rhs}\n")
                out.append(f"\t{ind}}}\n")
            if colored:
                out.append("\t}\n")
        elif op.code == ol.OP_RESTRICT:
            for f, r in zip(fields, rhs_names):
                out.append(f"\t{_field_name(op.dst, l - 1, mx, f, r)} = gen_restrictionForRes_{f}@{l} * "
                           f"{_field_name(op.src, l, mx, f, r)}\n")
        elif op.code == ol.OP_RESIDUAL_RESTRICT:
            for f, r in zip(fields, rhs_names):
                out.append(f"\t{_field_name(ol.BUF_RHS, l - 1, mx, f, r)} = gen_restrictionForRes_{f}@{l} * "
                           f"({_field_name(ol.BUF_RHS, l, mx, f, r)} - A@{l}*{_field_name(ol.BUF_SOL, l, mx, f, r)})\n")
        elif op.code == ol.OP_PROLONG_ADD:
            for f, r in zip(fields, rhs_names):
                out.append(f"\t{_field_name(ol.BUF_SOL, l, mx, f, r)} += {op.omega} * (gen_prolongationForSol_{f}@{l - 1} * "
                           f"{_field_name(op.src, l - 1, mx, f, r)})\n")
        elif op.code == ol.OP_PROLONG_SET:
            for f, r in zip(fields, rhs_names):
                out.append(f"\t{_field_name(op.dst, l, mx, f, r)} = gen_prolongationForSol_{f}@{l - 1} * "
                           f"{_field_name(op.src, l - 1, mx, f, r)}\n")
        elif op.code == ol.OP_COARSE_SOLVE:
            for f, r in zip(fields, rhs_names):
                out.append(f"\tgen_rhs_{f}@{l} = {r}@{l}\n")
                out.append(f"\tgen_error_{f}@{l} = 0\n")
            out.append(f"\t{cycle_name}@{l}()\n")
            for f, r in zip(fields, rhs_names):
                out.append(f"\tgen_error_{f}@{l} = gen_error_{f}@{l}\n")
        else:
            out.append(f"\t// op {ol.OP_NAMES.get(op.code, op.code)} on level {l}\n")
    out.append("}\n\n")
    return "".join(out)
