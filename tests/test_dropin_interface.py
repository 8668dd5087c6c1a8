"""The drop-in boundary of SURVEY.md 8(b): the B200 generator must accept the calls the reference's Optimizer makes.
Signatures are read from the reference SOURCE with `ast` when /root/reference is present (it is in the build
container; the GPU box has no copy -> skipped there); nothing of the reference is imported or executed."""
import ast
import inspect
import os

import pytest

from evostencils_b200 import problems
from evostencils_b200.program_generator import B200ProgramGenerator, B200ProgramGeneratorFAS

REF = "/root/reference/evostencils/code_generation"
METHODS = ("generate_storage", "initialize_code_generation", "generate_and_evaluate", "generate_cycle_function",
           "reinitialize")
READ_ATTRIBUTES = ("dimension", "finest_grid", "coarsening_factor", "min_level", "max_level", "equations", "operators",
                   "fields", "problem_name", "mpi_rank", "uses_FAS")


def _reference_signatures(filename, classname):
    tree = ast.parse(open(os.path.join(REF, filename)).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == classname)
    out = {}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef):
            args = [a.arg for a in fn.args.args][1:]
            n_default = len(fn.args.defaults)
            out[fn.name] = (args, len(args) - n_default, fn.args.vararg is not None)
    props = {fn.name for fn in cls.body if isinstance(fn, ast.FunctionDef)
             and any(isinstance(d, ast.Name) and d.id == "property" for d in fn.decorator_list)}
    return out, props


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present")
@pytest.mark.parametrize("filename,classname,ours", [("exastencils.py", "ProgramGenerator", B200ProgramGenerator),
                                                     ("exastencils_FAS.py", "ProgramGeneratorFAS", B200ProgramGeneratorFAS)])
def test_methods_accept_the_reference_call_signatures(filename, classname, ours):
    ref, ref_props = _reference_signatures(filename, classname)
    for name in METHODS:
        if name not in ref:
            continue
        ref_args, ref_required, ref_varargs = ref[name]
        sig = inspect.signature(getattr(ours, name))
        params = [p for p in list(sig.parameters.values())[1:]]
        if any(p.kind is inspect.Parameter.VAR_POSITIONAL for p in params) or ref_varargs:
            continue                                   # *args on either side: positional compatibility by construction
        names = [p.name for p in params]
        # every reference parameter exists under the same name at the same position (keyword and positional calls work)
        assert names[:len(ref_args)] == ref_args, (name, names, ref_args)
        required = sum(1 for p in params if p.default is inspect.Parameter.empty)
        assert required <= ref_required, (name, "requires more arguments than the reference")
    for attr in READ_ATTRIBUTES:
        if attr in ref_props:
            assert isinstance(inspect.getattr_static(ours, attr, None), property) or hasattr(ours, attr), attr


def test_constructor_raises_like_a_missing_compiler_without_a_gpu():
    from evostencils_b200 import backend
    if backend.load_library().evo_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(RuntimeError, match="Compiler not found"):
        B200ProgramGenerator(problem=problems.Poisson2D(3, 5))
