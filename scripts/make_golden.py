#!/usr/bin/env python
"""Generate tests/golden/*.json by running the REFERENCE's own Python half in this container.

Needs /root/reference (read-only) and is therefore only runnable here, not on the GPU box; its
outputs are committed.  DEAP is not installed, so a ~40-line stand-in of the three ``deap.gp`` classes
the reference's grammar touches is injected into ``sys.modules`` first; the
reference's real grammar (grammar/multigrid.py:409 ``generate_primitive_set``) and real text emitter
(code_generation/exastencils.py:318 ``generate_cycle_function``) then run unmodified on
OperatorInfo / EquationInfo objects built from our problem descriptors.

For every individual it records: the grammar string, the reference emitter's ExaSlang text, the op
list our lowering produces from the REFERENCE's tree, and (for small levels) the oracle's residual
history / fitness tuple for that op list.
"""
import json
import os
import random
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("EVOSTENCILS_REFERENCE", "/root/reference")


def install_deap_stub():
    gp = types.ModuleType("deap.gp")

    class Primitive:
        def __init__(self, name, args, ret):
            self.name, self.args, self.ret, self.arity = name, args, ret, len(args)

        def format(self, *a):
            return f"{self.name}({', '.join(a)})"

    class Terminal:
        def __init__(self, terminal, symbolic, ret):
            self.ret, self.value, self.arity = ret, terminal, 0
            self.name = str(terminal)
            self.conv_fct = str if symbolic else repr

        def format(self):
            return self.conv_fct(self.value)

    class PrimitiveSetTyped:
        def __init__(self, name, in_types, ret_type, prefix="ARG"):
            from collections import defaultdict
            self.terminals, self.primitives = defaultdict(list), defaultdict(list)
            self.name, self.ret, self.ins = name, ret_type, in_types
            self.mapping, self.context = {}, {"__builtins__": None}
            self.terms_count = self.prims_count = 0

        def addPrimitive(self, primitive, in_types, ret_type, name=None):
            name = name or primitive.__name__
            self._add(Primitive(name, in_types, ret_type))
            self.context[name] = primitive
            self.prims_count += 1

        def addTerminal(self, terminal, ret_type, name=None):
            symbolic = False
            if name is None and callable(terminal):
                name = terminal.__name__
            if name is not None:
                self.context[name] = terminal
                terminal = name
                symbolic = True
            prim = Terminal(terminal, symbolic, ret_type)
            self._add(prim)
            self.terms_count += 1

    gp.Primitive, gp.Terminal, gp.PrimitiveSetTyped = Primitive, Terminal, PrimitiveSetTyped
    deap = types.ModuleType("deap")
    deap.gp = gp
    sys.modules["deap"], sys.modules["deap.gp"] = deap, gp


def reference_problem_objects(problem):
    """OperatorInfo / EquationInfo lists as parser.extract_l2_information would deliver them."""
    import sympy
    from evostencils.grammar import multigrid as mg
    from evostencils.ir import base
    from evostencils.stencils import constant
    from evostencils_b200 import oplist as ol
    fields = [sympy.Symbol(f) for f in problem.fields]
    equations, operators = [], []
    for level in range(problem.min_level, problem.max_level + 1):
        table = problem.operator(level)
        for i, (eq, rhs) in enumerate(zip(problem.equation_names, problem.rhs_names)):
            terms = []
            for j, fld in enumerate(problem.fields):
                name = f"op{i}{j}"
                ent = [(ol.stencil_offset(p, problem.dim), float(table[i, j, p])) for p in range(27) if table[i, j, p] != 0]
                if not ent:
                    continue
                operators.append(mg.OperatorInfo(name, level, constant.Stencil(ent, problem.dim), base.Operator))
                terms.append(f"( {name}@{level} * {fld}@{level} )")
            e = mg.EquationInfo(eq, level, " + ".join(terms) + f" == {rhs}@{level}")
            e._associated_field = fields[i]
            equations.append(e)
        rw, pw = problem.restrict_weights(), problem.prolong_weights()
        for fld in problem.fields:
            operators.append(mg.OperatorInfo(f"gen_restrictionForRes_{fld}", level, constant.Stencil(
                [(ol.stencil_offset(p, problem.dim), float(rw[p])) for p in range(27) if rw[p] != 0], problem.dim),
                base.Restriction))
            operators.append(mg.OperatorInfo(f"gen_prolongationForSol_{fld}", level, constant.Stencil(
                [(ol.stencil_offset(p, problem.dim), float(pw[p])) for p in range(27) if pw[p] != 0], problem.dim),
                base.Prolongation))
    return equations, operators, fields


def reference_setup(problem, maximum_local_system_size=4):
    from evostencils.code_generation.exastencils import ProgramGenerator
    from evostencils.grammar import multigrid as mg
    from evostencils.ir import base, system
    equations, operators, fields = reference_problem_objects(problem)
    size = 2 ** problem.max_level
    finest = [base.Grid((size,) * problem.dim, (1.0 / size,) * problem.dim, problem.max_level) for _ in fields]
    cf = [(2,) * problem.dim for _ in fields]
    pg = object.__new__(ProgramGenerator)
    pg._cycle_name, pg._use_jacobi_prefix, pg._solver_cache = "gen_mgCycle", True, {}
    pg._dimension, pg._equations, pg._fields, pg._operators, pg._coarsening_factor = problem.dim, equations, fields, operators, cf

    def fresh_pset():
        approximation = system.Approximation("x", [base.Approximation(f.name, g) for f, g in zip(fields, finest)])
        rhs = system.RightHandSide("b", [base.RightHandSide(eq.rhs_name, g)
                                         for eq, g in zip([e for e in equations if e.level == problem.max_level], finest)])
        pset, _ = mg.generate_primitive_set(approximation, rhs, problem.dim, cf, problem.max_level, equations, operators,
                                            fields, maximum_local_system_size=maximum_local_system_size,
                                            depth=problem.max_level - problem.min_level)
        return pset
    storages = pg.generate_storage(problem.min_level, problem.max_level, finest)
    return pg, storages, fresh_pset


def grow_individual(pset, rng, size_limit=150):
    """The reference's genGrow (grammar/gp.py:6-52) driven by a seeded RNG."""
    import evostencils.grammar.gp as refgp
    refgp.random = rng
    expr = refgp.genGrow(pset, 0, 60, size_limit=size_limit)

    def fmt(nodes):
        # prefix list -> string (what deap's PrimitiveTree.__str__ does)
        string, stack = "", []
        for node in nodes:
            stack.append((node, []))
            while len(stack[-1][1]) == stack[-1][0].arity:
                prim, args = stack.pop()
                string = prim.format(*args)
                if not stack:
                    break
                stack[-1][1].append(string)
        return string
    return fmt(expr), len(expr)


def main():
    install_deap_stub()
    sys.path.insert(0, REF)
    from evostencils_b200 import fitness, lowering, oplist as ol, problems, tree
    from oracle import oracle as orc

    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    configs = [
        ("poisson2d", problems.Poisson2D(3, 6), 4, 10),
        ("elasticity2d", problems.LinearElasticity2D(3, 6), 4, 10),
        ("poisson3d", problems.Poisson3D(2, 5), 4, 6),
    ]
    for name, problem, max_lss, n_random in configs:
        pg, storages, fresh_pset = reference_setup(problem, max_lss)
        levels = problem.max_level - problem.min_level
        strings = [tree.v_cycle_individual(levels, 1, 1), tree.v_cycle_individual(levels, 2, 1, 23),
                   tree.v_cycle_individual(levels, 2, 2, 12, partitioning="single")]
        if problem.n_fields > 1:
            strings.append(tree.v_cycle_individual(levels, 1, 1, 14, smoother="decoupled_jacobi"))
        rng = random.Random(20261018)
        tries = 0
        while len(strings) < (4 if problem.n_fields > 1 else 3) + n_random and tries < 2000:
            tries += 1
            try:
                s, size = grow_individual(fresh_pset(), rng)
            except (RuntimeError, IndexError):
                continue
            strings.append(s)
        records = []
        for s in strings:
            pset = fresh_pset()
            ctx = dict(pset.context)
            ctx.pop("__builtins__", None)
            expression, _ = eval(s, {"__builtins__": {}}, ctx)
            prog = lowering.lower_cycle(expression, problem.min_level, problem.max_level, problem.n_fields, problem.dim,
                                        cgs_max_iters=problem.settings.cgs_max_iters, cgs_tol=problem.settings.cgs_tol)
            text = pg.generate_cycle_function(expression, storages, problem.min_level, problem.max_level, problem.max_level)
            # the tree must come out of the lowering untouched
            rec = {"individual": s, "exaslang": text, "program": prog.to_json()}
            oc = orc.OracleProblem(problem).build(prog)
            res = oc.solve(problem.settings.tol, problem.settings.max_iters, 1)
            t, cf, its = fitness.fitness_from_history(res.residuals, res.time_ms, problem.settings.max_iters)
            rec["oracle"] = {"iterations": res.iterations, "status": res.status,
                             "residuals": [float.hex(float(r)) for r in res.residuals],
                             "convergence_factor": cf, "fitness_iterations": its}
            records.append(rec)
            print(f"{name}: {len(prog.ops):3d} ops, iters={res.iterations:3d}, cf={cf:.6g}  {s[:70]}...")
        with open(os.path.join(out_dir, f"{name}.json"), "w") as f:
            json.dump({"problem": name, "min_level": problem.min_level, "max_level": problem.max_level,
                       "maximum_local_system_size": max_lss, "generator": "scripts/make_golden.py",
                       "records": records}, f, indent=1)


def main_fas():
    """FAS golden records: the reference's FAS grammar (generate_primitive_set(..., FAS=True)) and its FAS
    emitter (ProgramGeneratorFAS.generate_cycle_function, exastencils_FAS.py:428-439) on the shipped
    FAS_2D_Basic template (levels 10..8 of its knowledge file)."""
    sys.path.insert(0, REF)
    from evostencils.code_generation.exastencils_FAS import ProgramGeneratorFAS
    from evostencils.grammar import multigrid as mg
    from evostencils.ir import base, system
    from evostencils_b200 import fitness, lowering_fas, problems, tree
    from oracle import oracle as orc
    pg = ProgramGeneratorFAS("FAS_2D_Basic", "Solution", "RHS", "Residual", "Approximation", "RestrictionNode",
                             "CorrectionNode", "Laplace", "gamSten", "gen_mgCycle", "CGS",
                             build_path=os.path.join(REF, "example_problems"))
    depth = 2
    problem = problems.FAS2D(pg.max_level - depth, pg.max_level)

    def fresh_pset():
        approximation = system.Approximation("x", [base.Approximation(f.name, g) for f, g in zip(pg.fields, pg.finest_grid)])
        rhs = system.RightHandSide("b", [base.RightHandSide("RHS_u", g) for g in pg.finest_grid])
        pset, _ = mg.generate_primitive_set(approximation, rhs, pg.dimension, pg.coarsening_factor, pg.max_level,
                                            pg.equations, pg.operators, pg.fields, depth=depth, FAS=True)
        return pset
    strings = [tree.fas_v_cycle_individual(depth, 2, 2, 14, 1, "single"),
               tree.fas_v_cycle_individual(depth, 1, 1, 14, 2, "single"),
               tree.fas_v_cycle_individual(depth, 2, 1, 10, 0, "red_black"),
               tree.fas_v_cycle_individual(depth, 1, 2, 16, 3, "red_black")]
    rng = random.Random(4242)
    tries = 0
    while len(strings) < 9 and tries < 500:
        tries += 1
        try:
            s, size = grow_individual(fresh_pset(), rng, size_limit=60)
        except (RuntimeError, IndexError):
            continue
        strings.append(s)
    records = []
    ops_default = {l: problem.operator(l) for l in range(problem.min_level, problem.max_level + 1)}
    for n, s in enumerate(strings):
        ctx = dict(fresh_pset().context)
        ctx.pop("__builtins__", None)
        expression, _ = eval(s, {"__builtins__": {}}, ctx)
        text = pg.generate_cycle_function(expression)
        prog = lowering_fas.lower_fas_cycle(expression, problem.min_level, problem.max_level, 2,
                                            problem.settings.cgs_max_iters, problem.settings.damping,
                                            problem.restrict_weights(), problem.prolong_weights(), ops_default)
        rec = {"individual": s, "exaslang": text, "program": prog.to_json()}
        if n < 5:
            oc = orc.OracleProblem(problem).build(prog)
            res = oc.solve(problem.settings.tol, problem.settings.max_iters, 1)
            t, cf, its = fitness.fas_fitness(res.residuals, res.time_ms)
            rec["oracle"] = {"iterations": res.iterations, "status": res.status,
                             "residuals": [float.hex(float(r)) for r in res.residuals],
                             "convergence_factor": cf, "fitness_iterations": its}
            print(f"fas2d: {len(prog.ops):3d} ops, iters={res.iterations:3d}, cf={cf:.6g}  {s[:70]}...")
        else:
            print(f"fas2d: {len(prog.ops):3d} ops  {s[:90]}...")
        records.append(rec)
    with open(os.path.join(ROOT, "tests", "golden", "fas2d.json"), "w") as f:
        json.dump({"problem": "fas2d", "min_level": problem.min_level, "max_level": problem.max_level,
                   "generator": "scripts/make_golden.py", "records": records}, f, indent=1)


if __name__ == "__main__":
    if "--fas-only" not in sys.argv:
        main()
    else:
        install_deap_stub()
    main_fas()
