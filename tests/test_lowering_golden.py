"""Structure parity of the lowering against the reference's own emitter (CPU only).

tests/golden/*.json were produced by scripts/make_golden.py in the build container by running the
reference's real grammar and real ExaSlang emitter (exastencils.py:318 generate_cycle_function) on
trees from genGrow / hand-written V-cycles, and lowering the SAME reference trees with our lowering.
Here (no reference, no DEAP needed):
  1. our independent tree factory + lowering reproduces the recorded op lists exactly,
  2. our ExaSlang emitter's text equals the reference emitter's text statement by statement
     (operator sub-expression spellings and the local-system equation bodies are normalised away),
  3. the oracle reproduces the recorded residual histories bit for bit.
"""
import json
import os
import re

import numpy as np
import pytest

from evostencils_b200 import exaslang, fitness, lowering, oplist as ol, problems, tree

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PROBLEMS = {"poisson2d": problems.Poisson2D, "elasticity2d": problems.LinearElasticity2D, "poisson3d": problems.Poisson3D}


def load(name):
    with open(os.path.join(GOLDEN, f"{name}.json")) as f:
        data = json.load(f)
    prob = PROBLEMS[name](data["min_level"], data["max_level"])
    return prob, data["records"]


def all_records():
    out = []
    for name in PROBLEMS:
        prob, recs = load(name)
        out += [(name, i) for i in range(len(recs))]
    return out


def skeleton(text):
    """Statement skeleton of ExaSlang cycle text: operator products inside residual expressions and the
    equation bodies of local solves are dropped, everything else (targets, sources, weights, unknown
    lists, colouring, jacobi prefix, order) is kept."""
    lines = []
    for raw in text.splitlines():
        s = raw.strip()
        if not s:
            continue
        m = re.match(r"^(\S+@\[[-\d, ]+\]) => \(.*\) == (\S+@\[[-\d, ]+\])$", s)
        if m:
            lines.append(f"UNKNOWN {m.group(1)} RHS {m.group(2)}")
            continue
        m = re.match(r"^(gen_residual_\S+) = (\S+?)( - \(.*\))*$", s)
        if m and " - (" in s:
            lines.append(f"RESIDUAL {m.group(1)} {m.group(2)}")
            continue
        m = re.match(r"^(\S+) \+= (\S+) \* \((\S+) \* (\S+)\)$", s)
        if m:
            lines.append(f"CORRECT {m.group(1)} {float(m.group(2))!r} {m.group(3)} {m.group(4)}")
            continue
        m = re.match(r"^solve locally at (\S+) (with jacobi )?relax (\S+) \{$", s)
        if m:
            lines.append(f"SOLVE {m.group(1)} {'jacobi ' if m.group(2) else ''}{float(m.group(3))!r}")
            continue
        lines.append(s)
    return lines


@pytest.mark.parametrize("name,index", all_records())
def test_tree_factory_and_lowering_reproduce_reference_lowering(name, index):
    prob, recs = load(name)
    rec = recs[index]
    expression = tree.build_tree(prob, rec["individual"])
    prog = lowering.lower_cycle(expression, prob.min_level, prob.max_level, prob.n_fields, prob.dim,
                                cgs_max_iters=prob.settings.cgs_max_iters, cgs_tol=prob.settings.cgs_tol)
    golden = ol.Program.from_json(rec["program"])
    assert prog.structure() == golden.structure()
    assert sorted(prog.operators) == sorted(golden.operators)
    for l in prog.operators:
        np.testing.assert_array_equal(np.asarray(prog.operators[l]), np.asarray(golden.operators[l]))
    np.testing.assert_array_equal(prog.restrict_w, golden.restrict_w)
    np.testing.assert_array_equal(prog.prolong_w, golden.prolong_w)
    # lowering twice gives the same result: the tree is not mutated (the reference mutates .valid)
    again = lowering.lower_cycle(expression, prob.min_level, prob.max_level, prob.n_fields, prob.dim,
                                 cgs_max_iters=prob.settings.cgs_max_iters, cgs_tol=prob.settings.cgs_tol)
    assert again.structure() == prog.structure()


@pytest.mark.parametrize("name,index", all_records())
def test_emitted_text_matches_reference_emitter(name, index):
    prob, recs = load(name)
    rec = recs[index]
    golden = ol.Program.from_json(rec["program"])
    ours = exaslang.program_to_exaslang(golden, prob.fields, prob.rhs_names, prob.max_level)
    assert skeleton(ours) == skeleton(rec["exaslang"])


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_oracle_reproduces_golden_histories(oracle_mod, name):
    prob, recs = load(name)
    for rec in recs[:6]:
        prog = ol.Program.from_json(rec["program"])
        out = oracle_mod.OracleProblem(prob).build(prog).solve(prob.settings.tol, prob.settings.max_iters, 1)
        want = np.array([float.fromhex(h) for h in rec["oracle"]["residuals"]])
        assert out.iterations == rec["oracle"]["iterations"]
        assert np.array_equal(out.residuals, want, equal_nan=True)
        cf = fitness.fitness_from_history(out.residuals, out.time_ms, prob.settings.max_iters)[1]
        assert cf == rec["oracle"]["convergence_factor"]


def test_random_individuals_are_grammar_valid():
    import random
    for prob in (problems.Poisson2D(3, 7), problems.LinearElasticity2D(3, 6), problems.Poisson3D(2, 4)):
        rng = random.Random(0)
        for _ in range(25):
            s = tree.random_individual(prob, rng)
            expression = tree.build_tree(prob, s)
            prog = lowering.lower_cycle(expression, prob.min_level, prob.max_level, prob.n_fields, prob.dim)
            assert any(o.code == ol.OP_COARSE_SOLVE for o in prog.ops)      # guard types: CGS visited
            assert set(prog.operators) == set(range(prob.min_level, prob.max_level + 1))


def test_optimise_fuses_residual_and_restriction():
    prob = problems.Poisson2D(3, 6)
    expression = tree.build_tree(prob, tree.v_cycle_individual(3, 2, 1))
    prog = lowering.lower_cycle(expression, 3, 6, 1, 2)
    fused = lowering.optimise(prog)
    assert sum(o.code == ol.OP_RESIDUAL_RESTRICT for o in fused.ops) == 3
    assert not any(o.code == ol.OP_RESTRICT for o in fused.ops)
    assert len(fused.ops) == len(prog.ops) - 3
