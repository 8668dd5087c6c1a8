"""FAS: structure parity of the FAS lowering with the reference's FAS emitter (CPU), golden replay."""
import json
import os
import re

import numpy as np
import pytest

from evostencils_b200 import fitness, lowering_fas, oplist as ol, problems, tree

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fas2d.json")


def load():
    with open(GOLDEN) as f:
        data = json.load(f)
    return problems.FAS2D(data["min_level"], data["max_level"]), data["records"]


def parse_reference_fas_text(text):
    """Layer-4 text of ProgramGeneratorFAS (exastencils_FAS.py print_exa) -> statement skeleton."""
    out = []
    lines = [l.strip() for l in text.splitlines()]
    i = 0
    while i < len(lines):
        l = lines[i]
        m = re.match(r"loop over Solution@(\d+) \{", l)
        if l.startswith("color with"):
            # coloured in-place smoother: count the update statements until the closing brace of the colour block
            j, steps, level, w, newton = i + 1, 0, None, None, False
            while not (lines[j] == "}" and lines[j - 1].startswith("apply bc")):
                mm = re.match(r"Solution@(\d+) \+= (\S+) \* \(", lines[j])
                if mm:
                    steps += 1
                    level, w = int(mm.group(1)), float(mm.group(2))
                    newton = "exp(" in lines[j].split("/")[-1]
                j += 1
            out.append(("SMOOTH", level, "rb", "newton" if newton else "picard", steps, repr(w)))
            i = j + 1
            continue
        if m and i + 1 < len(lines) and lines[i + 1].startswith("Variable Solution_old"):
            level, j, steps, w, newton = int(m.group(1)), i + 2, 0, None, False
            while not lines[j].startswith("Solution<next>"):
                mm = re.match(r"Solution@(\d+) \+= (\S+) \* \(", lines[j])
                steps += 1
                w = float(mm.group(2))
                newton = "exp(" in lines[j].split("/")[-1]
                j += 1
            out.append(("SMOOTH", level, "jacobi", "newton" if newton else "picard", steps, repr(w)))
            i = j
            continue
        if m:
            nxt = lines[i + 1]
            mm = re.match(r"Solution@(\d+) -= Approximation@(\d+)", nxt)
            if mm:
                out.append(("FAS_SUB_APX", int(mm.group(1))))
            mm = re.match(r"Solution@(\d+) \+= \( (\S+) \* \( CorrectionNode@(\d+) \* Solution@(\d+) \) \)", nxt)
            if mm:
                out.append(("PROLONG_ADD", int(mm.group(1)), repr(float(mm.group(2)))))
        m = re.match(r"loop over Residual@(\d+) \{", l)
        if m:
            out.append(("RESIDUAL", int(m.group(1))))
        m = re.match(r"loop over Approximation@(\d+) \{", l)
        if m:
            out.append(("FAS_RESTRICT_SOL", int(m.group(1)) + 1))
        m = re.match(r"loop over RHS@(\d+) \{", l)
        if m:
            out.append(("FAS_COARSE_RHS", int(m.group(1)) + 1))
        m = re.match(r"CGS@(\d+) \(", l)
        if m:
            out.append(("COARSE_SOLVE", int(m.group(1))))
        i += 1
    return out


def program_skeleton(prog):
    out = []
    for o in prog.ops:
        name = ol.OP_NAMES[o.code]
        if o.code == ol.OP_SMOOTH:
            out.append(("SMOOTH", o.level, "rb" if o.mode == ol.MODE_REDBLACK else "jacobi",
                        "newton" if o.kind == ol.KIND_FAS_NEWTON else "picard", o.count, repr(float(o.omega))))
        elif o.code == ol.OP_PROLONG_ADD:
            out.append((name, o.level, repr(float(o.omega))))
        else:
            out.append((name, o.level))
    return out


def records():
    return list(range(len(load()[1])))


@pytest.mark.parametrize("index", records())
def test_fas_lowering_matches_reference_fas_emitter(index):
    prob, recs = load()
    rec = recs[index]
    golden = ol.Program.from_json(rec["program"])
    assert program_skeleton(golden) == parse_reference_fas_text(rec["exaslang"])
    # our own tree factory (FAS productions) + FAS lowering reproduce the recorded op list
    expression = tree.build_tree(prob, rec["individual"])
    prog = lowering_fas.lower_fas_cycle(expression, prob.min_level, prob.max_level, 2, prob.settings.cgs_max_iters,
                                        prob.settings.damping, prob.restrict_weights(), prob.prolong_weights(),
                                        {l: prob.operator(l) for l in range(prob.min_level, prob.max_level + 1)})
    assert prog.structure() == golden.structure()


def test_shipped_golden_text_of_the_reference():
    """example_problems/FAS_2D_Basic/FAS_2D_Basic.exa4:213-269 is a 3-level FAS V(2,2) with the template's
    Smoother function: our hand-built FAS V-cycle has the same statement order."""
    from evostencils_b200 import cycles
    prob = problems.FAS2D(3, 5)
    ops = cycles.fas_v_cycle(prob, 2, 2).ops
    names = [ol.OP_NAMES[o.code] for o in ops]
    assert names == (["SMOOTH"] * 2 + ["RESIDUAL", "FAS_RESTRICT_SOL", "FAS_COARSE_RHS"]) * 2 + ["COARSE_SOLVE"] + \
        (["FAS_SUB_APX", "PROLONG_ADD"] + ["SMOOTH"] * 2) * 2


def test_oracle_reproduces_fas_golden(oracle_mod):
    prob, recs = load()
    rec = next(r for r in recs if "oracle" in r and r["oracle"]["iterations"] < 120)
    prog = ol.Program.from_json(rec["program"])
    out = oracle_mod.OracleProblem(prob).build(prog).solve(prob.settings.tol, prob.settings.max_iters, 1)
    want = np.array([float.fromhex(h) for h in rec["oracle"]["residuals"]])
    assert out.iterations == rec["oracle"]["iterations"]
    assert np.array_equal(out.residuals, want, equal_nan=True)
