"""ExaSlang front-end (no Java): when the reference checkout is present, reading its problem files must
give exactly the hand-written descriptors of evostencils_b200.problems (CPU only; skipped on the GPU box)."""
import os

import numpy as np
import pytest

from evostencils_b200 import frontend, problems

REF = os.environ.get("EVOSTENCILS_REFERENCE", "/root/reference")
BASE = os.path.join(REF, "example_problems")
needs_ref = pytest.mark.skipif(not os.path.isdir(BASE), reason="reference checkout not available")

CASES = [
    ("Poisson/2D_FD_Poisson_fromL2.settings", "Poisson/2D_FD_Poisson_fromL2.knowledge", problems.Poisson2D),
    ("Poisson/3D_FD_Poisson_fromL2.settings", "Poisson/3D_FD_Poisson_fromL2.knowledge", problems.Poisson3D),
    ("LinearElasticity/2D_FD_LinearElasticity_fromL2.settings", "LinearElasticity/2D_FD_LinearElasticity_fromL2.knowledge",
     problems.LinearElasticity2D),
]


@needs_ref
@pytest.mark.parametrize("settings,knowledge,cls", CASES)
def test_files_equal_descriptors(settings, knowledge, cls):
    p = frontend.load_problem(BASE, settings, knowledge)
    q = cls()
    assert (p.dim, p.min_level, p.max_level) == (q.dim, q.min_level, q.max_level)
    assert p.fields == q.fields and p.rhs_names == q.rhs_names and p.equation_names == q.equation_names
    assert p.name == q.name
    for f in ("tol", "max_iters", "num_pre", "num_post", "damping", "red_black", "cgs_max_iters", "cgs_tol"):
        assert getattr(p.settings, f) == getattr(q.settings, f), f
    for level in (q.min_level, q.max_level):
        np.testing.assert_allclose(p.operator(level), q.operator(level), rtol=1e-15, atol=0)
    small = q.with_levels(q.min_level, min(q.max_level, 5))
    psmall = p.with_levels(small.min_level, small.max_level)
    for fi in range(q.n_fields):
        np.testing.assert_allclose(psmall.initial_solution(fi), small.initial_solution(fi), rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(psmall.rhs(fi), small.rhs(fi), rtol=1e-14, atol=1e-13)


@needs_ref
def test_knowledge_reader_matches_shipped_values():
    assert frontend.read_knowledge(os.path.join(BASE, "Helmholtz/2D_FD_Helmholtz_fromL3.knowledge")) == (2, 3, 7)
    assert frontend.read_knowledge(os.path.join(BASE, "FAS_2D_Basic/FAS_2D_Basic.knowledge")) == (2, 6, 10)


def test_solver_block_parser():
    s = frontend.read_solver_block("""generate solver for u in solEq with {
      solver_targetResReduction = 1e-6
      solver_maxNumIts = 17
      solver_smoother_jacobiType = false
      solver_smoother_numPre = 3
      solver_smoother_numPost = 3
      solver_smoother_damping = 0.8
      solver_smoother_coloring = "red-black"
      solver_cgs = "CG"
      solver_cgs_maxNumIts = 128
      solver_cgs_targetResReduction = 1e-3 }""")
    assert (s.tol, s.max_iters, s.num_pre, s.num_post, s.damping, s.red_black, s.cgs_max_iters, s.cgs_tol) == \
        (1e-6, 17, 3, 3, 0.8, True, 128, 1e-3)
