"""Host side of the population pipeline: the immutable grammar terminals are shared per problem object
(tree._problem_cache) and stencil tables / smoother statements are cached per operator object (lowering._per_object).
The caches must never serve a stale operator: run-time parameters (the wave number of the k / 2k / 4k Helmholtz runs,
exastencils.py:518-532) are part of the key, and the lowered programs must equal the uncached ones."""
import copy

import numpy as np

from evostencils_b200 import lowering, problems, tree


def test_terminals_are_shared_between_individuals():
    prob = problems.Poisson2D(5, 9)
    a = tree.system_operator(prob, 9, "A_0")
    b = tree.system_operator(prob, 9, "A_0")
    assert a is b
    c1 = tree.grammar_context(prob)
    c2 = tree.grammar_context(prob)
    assert c1["A_1"] is c2["A_1"] and c1["P_1"] is c2["P_1"] and c1["R_0"] is c2["R_0"]
    assert c1["u_and_f"][0] is not c2["u_and_f"][0]          # nodes the productions mutate stay fresh per context


def test_changed_parameters_get_their_own_operators():
    prob = problems.Helmholtz2D()
    a = tree.system_operator(prob, prob.max_level, "A_0")
    other = copy.copy(prob)                                   # what generate_and_evaluate does for the 2k / 4k runs
    other.parameters = dict(other.parameters)
    other.parameters["k"] = 2.0 * float(prob.parameters["k"])
    other.wave_number = complex(other.parameters["k"])
    b = tree.system_operator(other, other.max_level, "A_0")
    assert a is not b
    ta, tb = lowering.operator_table(a, prob.n_fields), lowering.operator_table(b, other.n_fields)
    assert not np.array_equal(ta, tb)
    assert np.array_equal(tb, other.operator(other.max_level))
    assert tree.system_operator(prob, prob.max_level, "A_0") is a


def _lower(prob, string):
    return lowering.optimise(lowering.lower_cycle(
        tree.build_tree(prob, string), prob.min_level, prob.max_level, prob.n_fields, prob.dim,
        cgs_max_iters=prob.settings.cgs_max_iters, cgs_tol=prob.settings.cgs_tol,
        default_restrict=prob.restrict_weights(), default_prolong=prob.prolong_weights()))


def test_cached_lowering_equals_uncached_lowering():
    import random
    for make in (lambda: problems.Poisson2D(5, 8), lambda: problems.LinearElasticity2D(4, 7), lambda: problems.Poisson3D(2, 5)):
        prob = make()
        rng = random.Random(11)
        strings = [tree.random_individual(prob, rng, maximum_local_system_size=4) for _ in range(6)]
        cached = [_lower(prob, s) for s in strings] + [_lower(prob, s) for s in strings]      # second pass: everything cached
        for k, s in enumerate(strings + strings):
            fresh_problem = make()                            # new problem object, cleared per-object cache: nothing shared
            lowering._PER_OBJECT.clear()
            ref = _lower(fresh_problem, s)
            got = cached[k]
            assert len(ref.ops) == len(got.ops)
            for x, y in zip(ref.ops, got.ops):
                assert (x.code, x.level, x.mode, x.kind, x.count, x.omega, x.dst, x.src, x.unknowns) == \
                       (y.code, y.level, y.mode, y.kind, y.count, y.omega, y.dst, y.src, y.unknowns)
            assert set(ref.operators) == set(got.operators)
            for l in ref.operators:
                assert np.array_equal(ref.operators[l], got.operators[l])
            assert np.array_equal(ref.restrict_w, got.restrict_w) and np.array_equal(ref.prolong_w, got.prolong_w)


def test_cached_tables_are_read_only():
    prob = problems.Poisson2D(5, 7)
    table = lowering.operator_table(tree.system_operator(prob, 7, "A_0"), 1)
    try:
        table[0, 0, 0] = 1.0
    except ValueError:
        return
    raise AssertionError("a cached stencil table must not be writable")
