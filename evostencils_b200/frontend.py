"""ExaSlang front-end without Java.

The reference obtains its problem description (operators per level, equations, fields) by running
the ExaStencils generator once and parsing the *debug layer-3 output* it writes
(reference: evostencils/code_generation/parser.py:25-111 ``extract_l2_information``), which needs
the Java tool chain.  This module reads the same information straight from the user's files:

* ``.knowledge``: dimensionality / minLevel / maxLevel (parser.py:114-125),
* ``.settings``: configName / basePathPrefix (parser.py:128-143),
* ``.exa2``: ``Domain``, ``Field`` (initial value, boundary expression, also the keyword-less layer-2
  short form the shipped Poisson 2D file uses), ``Operator ... from Stencil { [o] => expr }``,
  ``Equation``, ``Globals { Expr name = value }``,
* ``.exa3``: the ``generate solver for ... with { ... }`` block.

and turns them into a :class:`evostencils_b200.problems.Problem` whose coefficient tables are the
stencil expressions evaluated with ``vf_gridWidth_* = 2^-level`` (rediscretisation per level).
"""
from __future__ import annotations

import math
import os
import re
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import oplist as ol
from .problems import Problem, SolverSettings


def read_knowledge(path: str) -> Tuple[int, int, int]:
    dim = lo = hi = None
    with open(path) as f:
        for line in f:
            line = line.split("//")[0]
            tokens = line.split("=")
            if len(tokens) < 2:
                continue
            lhs = tokens[0].strip()
            if lhs == "dimensionality":
                dim = int(tokens[1].strip())
            elif lhs == "minLevel":
                lo = int(tokens[1].strip())
            elif lhs == "maxLevel":
                hi = int(tokens[1].strip())
    if dim is None or lo is None or hi is None:
        raise ValueError(f"{path}: dimensionality / minLevel / maxLevel missing")
    return dim, lo, hi


def read_settings(path: str) -> Dict[str, str]:
    out = {}
    with open(path) as f:
        for line in f:
            tokens = line.split("=")
            if len(tokens) >= 2:
                out[tokens[0].strip()] = tokens[1].strip().strip('"')
    return out


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def _sympify(expr: str, symbols: Dict[str, object]):
    import sympy
    from sympy.parsing.sympy_parser import parse_expr
    import keyword
    expr = expr.replace("^", "**")
    expr = re.sub(r"(\d+\.?\d*)j\b", r"(\1*I)", expr)           # ExaSlang complex literal 0.5j
    # ExaSlang identifiers may be Python keywords (the elasticity file has a global named `lambda`)
    renamed = {}
    for name in list(symbols):
        if keyword.iskeyword(name):
            renamed[name + "__"] = symbols[name]
            expr = re.sub(rf"\b{name}\b", name + "__", expr)
    symbols = {**symbols, **renamed}
    local = {"PI": sympy.pi, "I": sympy.I, "fabs": sympy.Abs, "max": sympy.Max, "min": sympy.Min,
             "sin": sympy.sin, "cos": sympy.cos, "exp": sympy.exp, "sqrt": sympy.sqrt}
    local.update(symbols)
    return parse_expr(expr, local_dict=local)


_CODE_CACHE: Dict[str, object] = {}      # compiled expressions (module level: problems stay picklable)


class ExaProblem(Problem):
    """A problem read from ExaSlang layer-2/3 files."""

    def __init__(self, name, dim, fields, rhs_names, equation_names, min_level, max_level, settings,
                 stencils, equations, globals_, boundary, rhs_expr):
        super().__init__(name=name, dim=dim, fields=tuple(fields), rhs_names=tuple(rhs_names),
                         equation_names=tuple(equation_names), min_level=min_level, max_level=max_level, settings=settings)
        self._stencils, self._equations, self._globals = stencils, equations, dict(globals_)
        self._boundary, self._rhs_expr = boundary, rhs_expr
        self.parameters = {k: float(v) for k, v in globals_.items() if _is_number(v)}

    def _symbols(self):
        import sympy
        syms = {f"vf_gridWidth_{a}": sympy.Symbol(f"vf_gridWidth_{a}") for a in "xyz"}
        for a in "xyz":
            for base in ("vf_nodePos", "vf_boundaryPos", "vf_boundaryCoord", "vf_nodePosition"):
                syms[f"{base}_{a}"] = sympy.Symbol(a)
        for k, v in self._globals.items():
            syms[k] = sympy.sympify(self.parameters.get(k, v))
        return syms

    def _namespace(self, extra=None):
        """Names an ExaSlang expression may use, for LITERAL evaluation (left to right, like the generated C++; a
        symbolic round trip would reorder the operands and change the last bit)."""
        import keyword
        ns = {"PI": math.pi, "sin": np.sin, "cos": np.cos, "exp": np.exp, "sqrt": np.sqrt, "fabs": np.abs,
              "max": np.maximum, "min": np.minimum, "__builtins__": {}}
        for k, v in self._globals.items():
            val = self.parameters.get(k, v)
            if not _is_number(val):
                val = eval(_pythonise(str(val)), dict(ns))      # noqa: S307 - a global defined by other globals
            ns[k + "__" if keyword.iskeyword(k) else k] = float(val) if not isinstance(val, complex) else val
        if extra:
            ns.update(extra)
        return ns

    def _operator_stencil(self, name: str, level: int) -> Dict[Tuple[int, ...], complex]:
        h = self.spacing(level)
        ns = self._namespace({f"vf_gridWidth_{a}": h for a in "xyz"})
        out = {}
        for offset, expr in self._stencils[name]:
            val = complex(eval(_pythonise(expr), ns))           # noqa: S307 - arithmetic of the user's problem file
            out[offset] = out.get(offset, 0) + val
        return out

    def operator(self, level: int) -> np.ndarray:
        import sympy
        nf = self.n_fields
        table = np.zeros((nf, nf, ol.STENCIL_POINTS), dtype=np.complex128)
        op_syms = {n: sympy.Symbol(n, commutative=False) for n in self._stencils}
        fld_syms = {f: sympy.Symbol(f, commutative=False) for f in self.fields}
        for i, eq_name in enumerate(self.equation_names):
            lhs = self._equations[eq_name][0]
            syms = dict(self._symbols()); syms.update(op_syms); syms.update(fld_syms)
            expr = sympy.expand(_sympify(lhs, syms))
            for term in sympy.Add.make_args(expr):
                coeff, factors = term.as_coeff_mul()
                ops = [f for f in factors if f in op_syms.values()]
                flds = [f for f in factors if f in fld_syms.values()]
                scal = [f for f in factors if f not in op_syms.values() and f not in fld_syms.values()]
                if len(flds) != 1 or len(ops) > 1:
                    raise ValueError(f"equation {eq_name}: unsupported term {term}")
                c = complex(sympy.Mul(coeff, *scal).evalf())
                j = self.fields.index(str(flds[0]))
                sten = self._operator_stencil(str(ops[0]), level) if ops else {(0,) * self.dim: 1.0}
                for off, v in sten.items():
                    table[i, j, ol.stencil_index(off)] += c * v
        return table.real.copy() if np.all(table.imag == 0) else table

    def _eval(self, key, expr, coords):
        code = _CODE_CACHE.get(expr)
        if code is None:
            code = _CODE_CACHE[expr] = compile(_pythonise(expr), "<exaslang>", "eval")
        extra = {}
        for a, c in zip("xyz", coords):
            for base in ("vf_nodePos", "vf_boundaryPos", "vf_boundaryCoord", "vf_nodePosition"):
                extra[f"{base}_{a}"] = c
        val = eval(code, self._namespace(extra))                 # noqa: S307 - arithmetic of the user's problem file
        return np.broadcast_to(np.asarray(val, dtype=np.float64), np.broadcast(*coords).shape).copy()

    def boundary_value(self, fi, level, *coords):
        expr = self._boundary.get(self.fields[fi])
        if isinstance(expr, dict):
            expr = expr.get("finest") if level == self.max_level else expr.get("coarser")
        if expr is None or _is_zero(expr):
            return None
        return self._eval(("bc", fi, level == self.max_level), expr, coords)

    def rhs_value(self, fi, level, *coords):
        expr = self._rhs_expr.get(self.rhs_names[fi])
        if expr is None or _is_zero(expr):
            return None
        return self._eval(("rhs", fi), expr, coords)


def _pythonise(expr: str) -> str:
    """ExaSlang arithmetic -> Python source: ``^`` is the power operator, identifiers that are Python keywords (the
    elasticity file has a global named ``lambda``) get a trailing ``__``."""
    import keyword
    expr = expr.replace("^", "**")
    return re.sub(r"[A-Za-z_]\w*", lambda m: m.group(0) + "__" if keyword.iskeyword(m.group(0)) else m.group(0), expr)


def _is_number(s) -> bool:
    try:
        float(s)
        return True
    except (TypeError, ValueError):
        return False


def _is_zero(expr: str) -> bool:
    return _is_number(expr) and float(expr) == 0.0


def read_solver_block(text: str) -> SolverSettings:
    s = SolverSettings()
    m = re.search(r"generate\s+solver\s+for.*?with\s*\{(.*?)\}", _strip_comments(text), flags=re.S)
    if not m:
        return s
    kv = dict(re.findall(r"(\w+)\s*=\s*([^\s]+)", m.group(1)))
    s.tol = float(kv.get("solver_targetResReduction", s.tol))
    s.max_iters = int(kv.get("solver_maxNumIts", s.max_iters))
    s.num_pre = int(kv.get("solver_smoother_numPre", s.num_pre))
    s.num_post = int(kv.get("solver_smoother_numPost", s.num_post))
    s.damping = float(kv.get("solver_smoother_damping", s.damping))
    s.red_black = kv.get("solver_smoother_coloring", '"red-black"').strip('"') == "red-black" and \
        kv.get("solver_smoother_jacobiType", "false") == "false"
    s.cgs_max_iters = int(kv.get("solver_cgs_maxNumIts", s.cgs_max_iters))
    s.cgs_tol = float(kv.get("solver_cgs_targetResReduction", s.cgs_tol))
    return s


def read_exa2(text: str, dim: int):
    """(fields, rhs fields, stencils, equations, globals, boundary expressions, rhs expressions)."""
    text = _strip_comments(text)
    stencils: Dict[str, List[Tuple[Tuple[int, ...], str]]] = {}
    for m in re.finditer(r"(?:Operator\s+)?(\w+)\s+from\s+Stencil\s*\{(.*?)\}", text, flags=re.S):
        entries = []
        for e in re.finditer(r"\[([^\]]+)\]\s*=>\s*([^\n]+)", m.group(2)):
            off = tuple(int(v) for v in e.group(1).split(","))
            entries.append((off, e.group(2).strip()))
        stencils[m.group(1)] = entries
    globals_: Dict[str, str] = {}
    for m in re.finditer(r"Globals\s*\{(.*?)\}", text, flags=re.S):
        for g in re.finditer(r"(?:Expr|Var|Val)\s+(\w+)\s*(?::\s*\w+)?\s*=\s*([^\n]+)", m.group(1)):
            globals_[g.group(1)] = g.group(2).strip()
    equations: Dict[str, Tuple[str, str]] = {}
    for m in re.finditer(r"(?:Equation\s+)?(\w+)\s*\{\s*([^{}]*?)==\s*([^{}]*?)\}", text, flags=re.S):
        if m.group(1) in ("Globals",):
            continue
        equations[m.group(1)] = (m.group(2).strip(), m.group(3).strip())
    decl, boundary = {}, {}
    for m in re.finditer(r"^\s*(?:Field\s+)?(\w+)(@\([^)]*\)|@\w+)?\s+with\s+\w+(?:<\w+>)?\s+on\s+Node\s+of\s+global(?:\s*=\s*([^\n]+))?",
                         text, flags=re.M):
        decl[m.group(1)] = (m.group(3) or "0.0").strip()
    for m in re.finditer(r"^\s*(?:Field\s+)?(\w+)(@\([^)]*\)|@\w+)?\s+on\s+boundary\s*=\s*([^\n]+)", text, flags=re.M):
        name, lvl, expr = m.group(1), m.group(2), m.group(3).strip()
        if lvl is None:
            boundary[name] = expr
        else:
            d = boundary.setdefault(name, {})
            if not isinstance(d, dict):
                d = boundary[name] = {"finest": d, "coarser": d}
            d["finest" if lvl == "@finest" else "coarser"] = expr
    rhs_names = [eq[1] for eq in equations.values()]
    fields = [n for n in decl if n not in rhs_names]
    rhs_expr = {n: decl[n] for n in rhs_names if n in decl}
    return fields, rhs_names, stencils, equations, globals_, boundary, rhs_expr


def load_problem(base_path: str, settings_path: str, knowledge_path: str) -> ExaProblem:
    """Problem from the reference's configuration triple (same arguments as ProgramGenerator.__init__,
    exastencils.py:40-41)."""
    settings = read_settings(os.path.join(base_path, settings_path))
    dim, lo, hi = read_knowledge(os.path.join(base_path, knowledge_path))
    name = settings["configName"]
    prefix = os.path.join(base_path, settings.get("basePathPrefix", "."))
    with open(os.path.join(prefix, f"{name}.exa2")) as f:
        fields, rhs_names, stencils, equations, globals_, boundary, rhs_expr = read_exa2(f.read(), dim)
    solver = SolverSettings()
    exa3 = os.path.join(prefix, f"{name}.exa3")
    if os.path.exists(exa3):
        with open(exa3) as f:
            solver = read_solver_block(f.read())
    # field order = sorted by name, equations sorted by associated field (parser.py:85, :98)
    fields = sorted(fields)
    eq_names = sorted(equations, key=lambda e: fields.index(equations[e][1].split("_")[-1]) if "_" in equations[e][1] else 0)
    rhs_sorted = [equations[e][1] for e in eq_names]
    return ExaProblem(name, dim, fields, rhs_sorted, eq_names, lo, hi, solver, stencils, equations, globals_, boundary, rhs_expr)
