"""TEST INFRASTRUCTURE: a minimal stand-in for the parts of DEAP the reference touches when its grammar is built and
its ``Optimizer`` is constructed (DEAP is not installed in this image).  Never imported by the
product package."""
import sys
import types
from collections import defaultdict


def install():
    if "deap" in sys.modules and getattr(sys.modules["deap"], "__evostencils_b200_stub__", False):
        return sys.modules["deap"]
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

    class _Context(dict):
        """optimization/program.py:920 pops the most recently added context entry before evaluating a grammar string.
        With a plain dict that entry is the coarse-grid-solver terminal ``CGS_n`` every individual needs; the notebook
        output (tutorial.ipynb:3373) shows the call worked against the DEAP of its time, so the stand-in keeps a
        sacrificial entry at the tail."""

        def __setitem__(self, key, value):
            dict.pop(self, "__stub_tail__", None)
            dict.__setitem__(self, key, value)
            dict.__setitem__(self, "__stub_tail__", None)

    class PrimitiveSetTyped:
        def __init__(self, name, in_types, ret_type, prefix="ARG"):
            self.terminals, self.primitives = defaultdict(list), defaultdict(list)
            self.name, self.ret, self.ins = name, ret_type, in_types
            self.mapping, self.context = {}, _Context()
            self.context["__builtins__"] = None
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
            self._add(Terminal(terminal, symbolic, ret_type))
            self.terms_count += 1

    class PrimitiveTree(list):
        def __str__(self):
            string, stack = "", []
            for node in self:
                stack.append((node, []))
                while len(stack[-1][1]) == stack[-1][0].arity:
                    prim, args = stack.pop()
                    string = prim.format(*args)
                    if not stack:
                        break
                    stack[-1][1].append(string)
            return string

    def compile_(expr, pset):
        return eval(str(expr), pset.context, {})    # noqa: S307 - grammar strings only

    def cx_one_point(a, b):
        return a, b

    gp.Primitive, gp.Terminal, gp.PrimitiveSetTyped, gp.PrimitiveTree = Primitive, Terminal, PrimitiveSetTyped, PrimitiveTree
    gp.compile, gp.cxOnePoint = compile_, cx_one_point

    base = types.ModuleType("deap.base")

    class Fitness:
        weights = ()

        def __init__(self, values=()):
            self.values = tuple(values)

        @property
        def valid(self):
            return len(self.values) != 0

    class Toolbox:
        def register(self, alias, function, *args, **kwargs):
            import functools
            setattr(self, alias, functools.partial(function, *args, **kwargs))

        def unregister(self, alias):
            delattr(self, alias)

    base.Fitness, base.Toolbox = Fitness, Toolbox

    creator = types.ModuleType("deap.creator")

    def create(name, base_cls, **kwargs):
        attrs = {k: v for k, v in kwargs.items() if not isinstance(v, type)}
        klass_attrs = {k: v for k, v in kwargs.items() if isinstance(v, type)}

        def __init__(self, *a, **kw):
            base_cls.__init__(self, *a, **kw)
            for k, v in klass_attrs.items():
                setattr(self, k, v())

        setattr(creator, name, type(name, (base_cls,), {**attrs, "__init__": __init__}))

    creator.create = create

    tools = types.ModuleType("deap.tools")
    tools.initIterate = lambda container, generator: container(generator())
    tools.initRepeat = lambda container, func, n: container(func() for _ in range(n))

    deap = types.ModuleType("deap")
    deap.__evostencils_b200_stub__ = True
    deap.gp, deap.base, deap.creator, deap.tools = gp, base, creator, tools
    for name, mod in (("deap", deap), ("deap.gp", gp), ("deap.base", base), ("deap.creator", creator), ("deap.tools", tools)):
        sys.modules[name] = mod
    return deap
