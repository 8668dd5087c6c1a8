"""Drop-in replacement for the evaluate path of the reference's ``ProgramGenerator``.

Reference: evostencils/code_generation/exastencils.py:39 (class), :485 ``generate_and_evaluate``, :318
``generate_cycle_function``, :586 ``generate_storage``, :445 ``initialize_code_generation``, :196
``reinitialize``.  ``Optimizer`` (optimization/program.py:68-72) accepts an instance of this class as
``program_generator`` unchanged: same attributes, same method names, same argument meaning, same
sentinel behaviour (``(infinity,)*3`` on failure, early un-averaged return on divergence).

What changes is what happens inside ``generate_and_evaluate``: instead of ExaSlang text -> Java code
generator (twice) -> ``make`` -> running the binary ``evaluation_samples`` times -> parsing stdout, the
tree is lowered to an op list and executed by the CUDA library (one CUDA-graph launch per sample).
"""
from __future__ import annotations

import math
import os
import re
import time
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import backend, exaslang, fitness, lowering, oplist as ol, problems
from .problems import Problem

JACOBI_COMPAT_MODES = ("intended", "exastencils_v1_1_noop")


# problems the reference defines on ExaSlang layer 3 / 4 only (no .exa2 to read): hand-written descriptors
_LAYER34_DESCRIPTORS = {"2D_FD_Helmholtz_fromL3": problems.Helmholtz2D, "FAS_2D_Basic": problems.FAS2D}
# the shipped layer-2 problems, used only when the configuration files are not on disk
_SHIPPED_DESCRIPTORS = {"2D_FD_Poisson_fromL2": problems.Poisson2D, "3D_FD_Poisson_fromL2": problems.Poisson3D,
                        "2D_FD_LinearElasticity_fromL2": problems.LinearElasticity2D, **_LAYER34_DESCRIPTORS}


def _problem_from_paths(settings_path: Optional[str], knowledge_path: Optional[str], base_path: Optional[str]) -> Problem:
    """The problem the reference's configuration triple describes (exastencils.py:39-110 + parser.py:25-143, which
    need the Java generator's debug output): the ``.settings`` file names the configuration, its ``.exa2`` / ``.exa3``
    files are read by :mod:`evostencils_b200.frontend` (equations, operators, boundary and right-hand-side
    expressions, globals, solver block), the ``.knowledge`` file gives dimensionality and levels (parser.py:114-125).
    Any user problem written on layer 2 works, not only the shipped ones; the two shipped problems that exist on
    layer 3 / 4 only (Helmholtz, FAS) use their descriptors."""
    from . import frontend
    base = base_path or ""
    sfile = os.path.join(base, settings_path) if settings_path else None
    kfile = os.path.join(base, knowledge_path) if knowledge_path else None
    if sfile and os.path.isfile(sfile):
        settings = frontend.read_settings(sfile)
        name = settings.get("configName")
        if not name:
            raise RuntimeError(f"{sfile}: configName missing")
        prefix = os.path.join(base, settings.get("basePathPrefix", "."))
        if os.path.isfile(os.path.join(prefix, f"{name}.exa2")):
            if not (kfile and os.path.isfile(kfile)):
                raise RuntimeError(f"knowledge file '{kfile}' not found")
            return frontend.load_problem(base, settings_path, knowledge_path)
        if name not in _LAYER34_DESCRIPTORS:
            raise RuntimeError(f"configuration '{name}' has no layer-2 file ({name}.exa2) and no descriptor")
        prob: Problem = _LAYER34_DESCRIPTORS[name]()
    else:
        stem = os.path.basename(settings_path or knowledge_path or "").split(".")[0].strip()
        if stem not in _SHIPPED_DESCRIPTORS:
            raise RuntimeError(f"settings file '{sfile}' not found and '{stem}' is not a shipped configuration")
        prob = _SHIPPED_DESCRIPTORS[stem]()
    if kfile and os.path.isfile(kfile):
        dim, lo, hi = frontend.read_knowledge(kfile)
        if dim != prob.dim:
            raise RuntimeError("dimensionality of the knowledge file does not match the problem")
        prob = prob.with_levels(lo, hi)
    return prob


class _FastGilSwitch:
    """While trees are lowered by background threads (pure Python) the submitting thread returns from a C call every
    few hundred microseconds and must get the interpreter lock back quickly: with the default switch interval of 5 ms it
    would wait up to 5 ms per call and the device would run dry.  Process-wide setting, reference counted."""
    _lock = None
    _users = 0
    _saved = None

    def __enter__(self):
        import sys
        import threading
        cls = _FastGilSwitch
        if cls._lock is None:
            cls._lock = threading.Lock()
        with cls._lock:
            if cls._users == 0:
                cls._saved = sys.getswitchinterval()
                sys.setswitchinterval(min(cls._saved, 2e-4))
            cls._users += 1
        return self

    def __exit__(self, *exc):
        import sys
        cls = _FastGilSwitch
        with cls._lock:
            cls._users -= 1
            if cls._users == 0 and cls._saved is not None:
                sys.setswitchinterval(cls._saved)
        return False


class _LoweredStream:
    """Programs of a list of trees, lowered by a background thread in order (len() and iteration like a list; a tree
    the lowering rejects yields None)."""

    def __init__(self, generator, expressions, min_level):
        import queue
        import threading
        self._n = len(expressions)
        self._q = queue.Queue(maxsize=256)

        def produce():
            for e in expressions:
                try:
                    self._q.put(generator._finalise(generator.lower(e, min_level)))
                except lowering.LoweringError:
                    self._q.put(None)
                except BaseException as exc:     # surfaced in the consumer
                    self._q.put(exc)
                    return

        self._thread = threading.Thread(target=produce, daemon=True)
        self._thread.start()

    def __len__(self):
        return self._n

    def __iter__(self):
        for _ in range(self._n):
            item = self._q.get()
            if isinstance(item, BaseException):
                raise item
            yield item


class B200ProgramGenerator:
    def __init__(self, absolute_compiler_path: Optional[str] = None, base_path: Optional[str] = None,
                 settings_path: Optional[str] = None, knowledge_path: Optional[str] = None,
                 platform_path: Optional[str] = None, mpi_rank: int = 0, solution_equations=None,
                 cycle_name: str = "gen_mgCycle", model_based_estimation: bool = False, evaluation_timeout=300,
                 code_generation_timeout=300, c_compiler_timeout=120, solver_iteration_limit=None, *,
                 problem: Optional[Problem] = None, device: Optional[int] = None,
                 jacobi_compat: str = "intended", fuse: bool = True):
        if jacobi_compat not in JACOBI_COMPAT_MODES:
            raise ValueError(f"jacobi_compat must be one of {JACOBI_COMPAT_MODES}")
        self._average_generation_time = 0          # written by Optimizer (program.py:850-851)
        self._counter = 0
        self.timeout_evaluate = evaluation_timeout
        self.timeout_exastencils_compiler = code_generation_timeout
        self.timeout_c_compiler = c_compiler_timeout
        self._solver_iteration_limit = solver_iteration_limit
        self._absolute_compiler_path = absolute_compiler_path
        self._base_path = base_path
        self._settings_path, self._knowledge_path, self._platform_path = settings_path, knowledge_path, platform_path
        self._mpi_rank = mpi_rank
        self._cycle_name = cycle_name
        self._use_jacobi_prefix = not model_based_estimation      # exastencils.py:64-70
        self._solution_equations = solution_equations
        self.jacobi_compat = jacobi_compat
        self.fuse = fuse
        self._problem = problem if problem is not None else _problem_from_paths(settings_path, knowledge_path, base_path)
        self._original_min_level, self._original_max_level = self._problem.min_level, self._problem.max_level
        # the reference raises RuntimeError("Compiler not found. Aborting.") when its tool chain is missing
        # (exastencils.py:104-108); ours is the CUDA library + a device.  No CPU fallback.
        try:
            lib = backend.load_library()
            ndev = lib.evo_device_count()
        except backend.BackendError as e:
            raise RuntimeError(f"Compiler not found. Aborting. ({e})")
        if ndev <= 0:
            raise RuntimeError("Compiler not found. Aborting. (no CUDA device visible; the B200 backend has no CPU fallback)")
        self._device = (mpi_rank % ndev) if device is None else device
        self._compiler_available = True
        self._device_problems: Dict[Tuple, backend.DeviceProblem] = {}
        self._cycle_registry: Dict[int, dict] = {}
        self._l2_cache = None
        self.last_outcome = None
        self.total_kernel_launches = 0

    # ---- attributes read by Optimizer / scripts (scripts/optimize.py:60-68, program.py:116-124, :808) ----
    @property
    def uses_FAS(self):
        return False

    @property
    def absolute_compiler_path(self):
        return self._absolute_compiler_path

    @property
    def knowledge_path(self):
        return self._knowledge_path

    @property
    def settings_path(self):
        return self._settings_path

    @property
    def problem_name(self):
        return self._problem.name

    @property
    def compiler_available(self):
        return self._compiler_available

    @property
    def base_path(self):
        return self._base_path

    @property
    def platform_path(self):
        return self._platform_path

    @property
    def dimension(self):
        return self._problem.dim

    @property
    def min_level(self):
        return self._problem.min_level

    @property
    def max_level(self):
        return self._problem.max_level

    @property
    def mpi_rank(self):
        return self._mpi_rank

    @property
    def solution_equations(self):
        return self._solution_equations

    @property
    def solver_iteration_limit(self):
        return self._solver_iteration_limit

    @property
    def problem(self) -> Problem:
        return self._problem

    @property
    def coarsening_factor(self):
        return [tuple([2] * self.dimension) for _ in self._problem.fields]

    def _l2(self):
        """equations / operators / fields / finest_grid in the classes ``generate_primitive_set`` expects
        (grammar/multigrid.py:15-71): the reference's own classes when the package is importable."""
        key = (self._problem.min_level, self._problem.max_level, tuple(sorted(self._problem.parameters.items())))
        if self._l2_cache is not None and self._l2_cache[0] == key:
            return self._l2_cache[1]
        p = self._problem
        try:
            import sympy
            from evostencils.grammar import multigrid as mg      # needs DEAP (grammar/gp.py:3)
            from evostencils.ir import base
            from evostencils.stencils import constant
            fields = [sympy.Symbol(f) for f in p.fields]
            mk_op = lambda name, level, ent, typ: mg.OperatorInfo(name, level, constant.Stencil(ent, p.dim), typ)
            types = (base.Operator, base.Restriction, base.Prolongation)
            mk_eq = mg.EquationInfo
            mk_grid = base.Grid
        except Exception:
            fields = [SimpleNamespace(name=f) for f in p.fields]
            mk_op = lambda name, level, ent, typ: SimpleNamespace(name=name, level=level, operator_type=typ,
                                                                  stencil=SimpleNamespace(entries=tuple(ent), dimension=p.dim))
            types = ("Operator", "Restriction", "Prolongation")
            mk_eq = lambda name, level, expr: SimpleNamespace(name=name, level=level, rhs_name=expr.split("==")[1].strip().split("@")[0],
                                                              _associated_field=None)
            mk_grid = lambda size, spacing, level: SimpleNamespace(size=size, spacing=spacing, level=level, dimension=len(size))
        equations, operators = [], []
        for level in range(p.min_level, p.max_level + 1):
            table = p.operator(level)
            for i, (eq, rhs) in enumerate(zip(p.equation_names, p.rhs_names)):
                terms = []
                for j, fld in enumerate(p.fields):
                    ent = [(ol.stencil_offset(q, p.dim), complex(table[i, j, q]) if np.iscomplexobj(table) else float(table[i, j, q]))
                           for q in range(ol.STENCIL_POINTS) if table[i, j, q] != 0]
                    if not ent:
                        continue
                    name = f"A{i}{j}"
                    operators.append(mk_op(name, level, ent, types[0]))
                    terms.append(f"( {name}@{level} * {fld}@{level} )")
                e = mk_eq(eq, level, " + ".join(terms) + f" == {rhs}@{level}")
                e._associated_field = fields[i]
                equations.append(e)
            rw, pw = p.restrict_weights(), p.prolong_weights()
            for fld in p.fields:
                operators.append(mk_op(f"gen_restrictionForRes_{fld}", level,
                                       [(ol.stencil_offset(q, p.dim), float(rw[q])) for q in range(27) if rw[q] != 0], types[1]))
                operators.append(mk_op(f"gen_prolongationForSol_{fld}", level,
                                       [(ol.stencil_offset(q, p.dim), float(pw[q])) for q in range(27) if pw[q] != 0], types[2]))
        size = 2 ** p.max_level
        finest = [mk_grid(tuple([size] * p.dim), tuple([1.0 / size] * p.dim), p.max_level) for _ in p.fields]
        self._l2_cache = (key, (equations, operators, fields, finest))
        return self._l2_cache[1]

    @property
    def equations(self):
        return self._l2()[0]

    @property
    def operators(self):
        return self._l2()[1]

    @property
    def fields(self):
        return self._l2()[2]

    @property
    def finest_grid(self):
        return self._l2()[3]

    # ---- methods called by Optimizer -------------------------------------------------------------------
    def generate_storage(self, min_level: int, max_level: int, finest_grids=None) -> list:
        """Opaque per-level handles (the reference returns CycleStorage objects, exastencils.py:586-592;
        callers only pass them back)."""
        return [SimpleNamespace(level=l) for l in range(max_level, min_level - 1, -1)]

    def _device_problem(self, min_level: int, max_level: int, problem: Optional[Problem] = None) -> backend.DeviceProblem:
        p = (problem or self._problem).with_levels(min_level, max_level)
        key = (min_level, max_level, tuple(sorted(p.parameters.items())))
        if key not in self._device_problems:
            if len(self._device_problems) >= 4:                 # keep HBM use bounded across generalisation steps
                k0 = next(iter(self._device_problems))
                self._device_problems.pop(k0).close()
            self._device_problems[key] = backend.DeviceProblem(p, device=self._device)
        return self._device_problems[key]

    def initialize_code_generation(self, min_level: int, max_level: int):
        """One-time set-up per level range (the reference runs the Java generator here, :445-466)."""
        start = time.time()
        self._device_problem(min_level, max_level)
        if self._counter == 0:
            self._counter += 1
            self._average_generation_time += (time.time() - start - self._average_generation_time) / self._counter
        return f"b200://{self.problem_name}/{min_level}-{max_level}"

    def reinitialize(self, min_level, max_level, global_expressions=None):
        """New level range / PDE parameters (generalisation step, program.py:110-146 -> exastencils.py:196-215)."""
        import copy
        p = copy.copy(self._problem).with_levels(min_level, max_level)
        if global_expressions:
            p.parameters = dict(p.parameters)
            for k, v in global_expressions.items():
                if k in p.parameters:
                    p.parameters[k] = float(v)
            if "k" in global_expressions and hasattr(p, "wave_number"):
                p.wave_number = complex(float(global_expressions["k"]))
        self._problem = p
        self._l2_cache = None

    def lower(self, expression, min_level: int, max_level: Optional[int] = None) -> ol.Program:
        p = self._problem
        prog = lowering.lower_cycle(expression, min_level, self.max_level if max_level is None else max_level,
                                    p.n_fields, p.dim, use_jacobi_prefix=self._use_jacobi_prefix,
                                    cgs_max_iters=p.settings.cgs_max_iters, cgs_tol=p.settings.cgs_tol,
                                    cycle_registry=self._cycle_registry,
                                    default_restrict=p.restrict_weights(), default_prolong=p.prolong_weights())
        return prog

    def _finalise(self, prog: ol.Program) -> ol.Program:
        prog = lowering.apply_jacobi_compat(prog, self.jacobi_compat)
        if self.fuse:
            prog = lowering.optimise(prog)
        return prog

    def generate_cycle_function(self, expression, storages, min_level: int, level: int, max_level: int,
                                use_global_weights: bool = False) -> str:
        prog = self.lower(expression, min_level, max_level)
        # a later run on finer levels may call this cycle as its coarse-grid solver (multi-run mode,
        # exastencils.py:893-896): remember the statements under the level it was generated for
        self._cycle_registry[level] = {"ops": list(prog.ops), "operators": dict(prog.operators)}
        return exaslang.program_to_exaslang(prog, self._problem.fields, self._problem.rhs_names, level, self._cycle_name)

    def _apply_sentinels(self, time_ms, cf, iters, infinity):
        """evaluate(), exastencils.py:435-443, for identical samples (cf / iterations are deterministic)."""
        if iters >= infinity or cf > 1:
            return time_ms, cf, iters
        if math.isinf(cf) or math.isnan(cf):
            return infinity, infinity, infinity
        return time_ms, cf, iters

    def _timeout_ms(self) -> int:
        """``evaluation_timeout`` (seconds, exastencils.py:42) as the watchdog of the device-side solver loop."""
        t = self.timeout_evaluate
        return int(min(max(float(t), 0.0) * 1000.0, 2 ** 31 - 1)) if t else 0

    def _evaluate_program(self, prog: ol.Program, min_level, problem, infinity, evaluation_samples):
        dev = self._device_problem(min_level, self.max_level, problem)
        s = dev.problem.settings
        helmholtz = dev.problem.kind == ol.PROBLEM_HELMHOLTZ
        if problem is not None:
            # PDE parameters patched for this run (exastencils.py:269-288 rewrites the generated globals):
            # the operators are rediscretised with the new values
            import copy
            prog = copy.copy(prog)
            prog.operators = {l: dev.problem.operator(l) for l in prog.operators}
        cyc = dev.build(prog)
        try:
            if helmholtz:
                out = cyc.helmholtz_solve(s.tol, s.max_iters, samples=max(1, int(evaluation_samples)), timeout_ms=self._timeout_ms())
            else:
                out = cyc.solve(s.tol, s.max_iters, samples=max(1, int(evaluation_samples)), timeout_ms=self._timeout_ms())
        finally:
            cyc.close()
        self.last_outcome = out
        self.total_kernel_launches += out.kernel_launches * max(1, int(evaluation_samples))
        if out.status == 2:                                     # evaluation timed out (:430-433, :476-483)
            return infinity, infinity, infinity
        if helmholtz:
            t, cf, its = fitness.helmholtz_fitness(out.residuals, out.time_ms, s.max_iters, infinity,
                                                   self._solver_iteration_limit, s.tol)
        else:
            t, cf, its = fitness.fitness_from_history(out.residuals, out.time_ms, s.max_iters, infinity,
                                                      self._solver_iteration_limit)
        return self._apply_sentinels(t, cf, its, infinity)

    def generate_and_evaluate(self, expression, storages, min_level: int, max_level: int, solver_program: str,
                              infinity=1e100, evaluation_samples=3, global_variable_values=None):
        """Same contract as the reference (exastencils.py:485-537): never raises for a *bad individual* -- a tree that
        cannot be lowered ("code generation failed", :499-510), an op list the library rejects as malformed, a
        diverging or timed-out solve all give ``(infinity,)*3``.  Failures of the infrastructure (no device, CUDA
        error, out of memory, a statement the library does not implement) are NOT fitness values and propagate as
        ``backend.BackendError``."""
        if global_variable_values is None:
            global_variable_values = {}
        start = time.time()
        try:
            prog = self._finalise(self.lower(expression, min_level))
        except lowering.LoweringError:
            return infinity, infinity, infinity                 # "Code generation failed" (:499-510)
        self._counter += 1
        self._average_generation_time += (time.time() - start - self._average_generation_time) / self._counter
        mapping = dict(global_variable_values)
        n = 3 if "k" in mapping else 1                          # :518-521
        avg = [0.0, 0.0, 0.0]
        for _ in range(n):
            problem = None
            if mapping:
                import copy
                problem = copy.copy(self._problem)
                problem.parameters = dict(problem.parameters)
                for k, v in mapping.items():
                    if k in problem.parameters:
                        problem.parameters[k] = float(v)
                if "k" in mapping:
                    problem.wave_number = complex(float(mapping["k"]))
                    problem.parameters["k"] = float(mapping["k"])
            try:
                t, cf, its = self._evaluate_program(prog, min_level, problem, infinity, evaluation_samples)
            except backend.BackendError as e:
                if e.infrastructure:
                    raise
                t, cf, its = infinity, infinity, infinity
            avg[0] += t; avg[1] += cf; avg[2] += its
            if its >= infinity or cf > 1:                       # :529-530
                return tuple(avg)
            if n > 1:
                mapping["k"] *= 2
        return avg[0] / n, avg[1] / n, avg[2] / n

    # ---- beyond the reference surface: a whole generation in one call -----------------------------------
    def evaluate_population(self, expressions: Sequence, min_level: Optional[int] = None, infinity=1e100,
                            evaluation_samples: int = 1, max_in_flight: int = 48, programs: Optional[Sequence[ol.Program]] = None,
                            solo_timing: bool = True, keep_for_retime: bool = False):
        """Fitness tuples of many individuals; a sliding window of ``max_in_flight`` solves runs concurrently on the
        GPU, one CUDA stream and one device-side solver loop each (the reference evaluates one after the other,
        program.py:491), while the host lowers and builds the next individuals (trees are lowered by a background
        thread: the main thread spends its time inside C calls that release the interpreter lock).  ~48 in flight
        saturate a B200 (the kernel launch rate, not the SMs, bounds this regime; more in flight is slower).
        Returns (list of tuples, milliseconds of the whole pipeline).

        ``solo_timing`` (default): the time entry of every converged individual is measured again with the GPU to
        itself (a few iterations, extrapolated to its iteration count), so that it is the same objective
        ``generate_and_evaluate`` returns and does not depend on the batch composition.  With ``False`` the time is the
        individual's span inside the concurrent batch -- higher throughput, but NOT comparable between batches.
        ``keep_for_retime``: return the concurrent-pipeline results but keep the finished cycles, so that
        :meth:`finish_retime` can re-time them later when the device is idle (several generators -- e.g. the two problems
        of a generation -- run their pipelines at the same time from different host threads, then re-time in turn).
        Helmholtz problems (outer BiCGStab) are evaluated one after the other through the same path as
        ``generate_and_evaluate``."""
        min_level = self.min_level if min_level is None else min_level
        dev = self._device_problem(min_level, self.max_level)
        s = dev.problem.settings
        results: List[Tuple[float, float, float]] = []
        total_ms = 0.0
        progs: List[Optional[ol.Program]] = []
        if programs is not None:
            progs = [self._finalise(p) for p in programs]
        else:
            # a sized iterable is consumed lazily by the lowering thread (e.g. trees built on the fly from grammar strings)
            progs = _LoweredStream(self, expressions if hasattr(expressions, "__len__") else list(expressions), min_level)
        sentinel = (infinity, infinity, infinity)
        if dev.problem.kind == ol.PROBLEM_HELMHOLTZ:
            for p in progs:
                if p is None:
                    results.append(sentinel)
                    continue
                try:
                    results.append(self._evaluate_program(p, min_level, None, infinity, evaluation_samples))
                    total_ms += self.last_outcome.time_ms * max(1, evaluation_samples)
                except backend.BackendError as e:
                    if e.infrastructure:
                        raise
                    results.append(sentinel)
            return results, total_ms
        # ---- pipeline: a sliding window of `max_in_flight` solves; each one is enqueued as soon as its cycle is built
        # (evo_cycle_solve_begin) and collected in order (evo_cycle_solve_end) while the host builds the next ones
        from collections import deque
        window: deque = deque()
        done: List[Tuple[int, backend.DeviceCycle, backend.SolveOutcome]] = []
        slots: List[Optional[Tuple[float, float, float]]] = [None] * len(progs)
        t_start = time.perf_counter()

        def tuple_of(o):
            if o.status == 2:
                return sentinel
            t, cf, its = fitness.fitness_from_history(o.residuals, o.time_ms, s.max_iters, infinity, self._solver_iteration_limit)
            return self._apply_sentinels(t, cf, its, infinity)

        kept: List[Tuple[int, backend.DeviceCycle]] = []

        def finish(batch):
            # solo_timing: the device is idle here -> contention-free time of the members that converged
            for j, c, o in batch:
                keep = False
                try:
                    if solo_timing:
                        o = c.solve_retime()
                    self.total_kernel_launches += o.kernel_launches
                    slots[j] = tuple_of(o)
                    keep = keep_for_retime
                finally:
                    if keep:
                        kept.append((j, c))
                    else:
                        c.close()

        gil = _FastGilSwitch() if programs is None else None
        if gil:
            gil.__enter__()
        try:
            for j, p in enumerate(progs):
                if p is None:
                    slots[j] = sentinel
                    continue
                try:
                    c = dev.build(p)
                except backend.BackendError as e:
                    if e.infrastructure:
                        raise
                    slots[j] = sentinel
                    continue
                try:
                    c.solve_begin(s.tol, s.max_iters, self._timeout_ms())
                except backend.BackendError as e:
                    c.close()
                    if e.infrastructure:
                        raise
                    slots[j] = sentinel
                    continue
                window.append((j, c))
                if len(window) >= max_in_flight:
                    k, ck = window.popleft()
                    done.append((k, ck, ck.solve_end()))
                if len(done) >= 512:                   # bound the device memory held by finished cycles
                    while window:
                        k, ck = window.popleft()
                        done.append((k, ck, ck.solve_end()))
                    finish(done)
                    done = []
            while window:
                k, ck = window.popleft()
                done.append((k, ck, ck.solve_end()))
            finish(done)
            done = []
        finally:
            if gil:
                gil.__exit__(None, None, None)
            for _, c in window:
                c.close()
            for _, c, _o in done:
                c.close()
        total_ms += (time.perf_counter() - t_start) * 1e3
        results = [r if r is not None else sentinel for r in slots]
        if keep_for_retime:
            self._pending_retime = (kept, results, tuple_of)
        return results, total_ms

    def finish_retime(self):
        """Second phase of ``evaluate_population(..., keep_for_retime=True)``: with the device idle, measure the time
        objective of every converged member again with the GPU to itself; returns (fitness tuples, milliseconds)."""
        kept, results, tuple_of = getattr(self, "_pending_retime", ([], [], None))
        self._pending_retime = ([], [], None)
        t0 = time.perf_counter()
        try:
            for j, c in kept:
                results[j] = tuple_of(c.solve_retime())
        finally:
            for _, c in kept:
                c.close()
        return results, (time.perf_counter() - t0) * 1e3

    def close(self):
        for d in self._device_problems.values():
            d.close()
        self._device_problems.clear()


# the name a user of the reference would look for
ProgramGenerator = B200ProgramGenerator


class B200ProgramGeneratorFAS:
    """Drop-in for ``ProgramGeneratorFAS`` (reference: code_generation/exastencils_FAS.py:11-445): same
    constructor arguments, ``uses_FAS``, ``generate_and_evaluate(*args, **kwargs)`` picking the ``Cycle`` out
    of its positional arguments (:396-404), ``generate_cycle_function(*args)``, the dummy
    ``generate_storage`` / ``initialize_code_generation`` (:441-445).

    Fitness follows the FAS ``parse_output`` (:370-394): ``c = (res_final/res_initial)^(1/n)`` from the
    4-digit prints of the template's Solve loop, ``n`` = iterations, ``(1e100,)*3`` on abort / non-finite c.
    ``cumulative_timer=True`` reproduces the template's time accounting (``t_sol +=
    getTotalFromTimer('cycle')`` after every iteration, FAS_2D_Basic_template.exa4:148-151: the timer is
    cumulative, so t_sol = sum_k k * t_cycle); set it to False for the plain solve time."""

    def __init__(self, problem_name="FAS_2D_Basic", solution="Solution", rhs="RHS", residual="Residual",
                 FASApproximation="Approximation", restriction="RestrictionNode", prolongation="CorrectionNode",
                 op_linear="Laplace", op_nonlinear="gamSten", fct_name_mgcycle="gen_mgCycle", fct_cgs="CGS",
                 fct_smoother=None, mpi_rank=0, platform_file=None, build_path=None, exastencils_compiler=None, *,
                 problem: Optional[Problem] = None, device: Optional[int] = None, cumulative_timer: bool = True):
        self.problem_name = problem_name
        self.mpi_rank = mpi_rank
        self.build_path = build_path
        self.cumulative_timer = cumulative_timer
        self._cycle_name = fct_name_mgcycle
        if problem is None:
            problem = problems.FAS2D()
            if build_path:
                kpath = os.path.join(build_path, problem_name, f"{problem_name}.knowledge")
                if os.path.exists(kpath):
                    from .frontend import read_knowledge
                    dim, lo, hi = read_knowledge(kpath)
                    problem = problem.with_levels(lo, hi)
        self._problem = problem
        self.min_level, self.max_level, self.dimension = problem.min_level, problem.max_level, problem.dim
        try:
            lib = backend.load_library()
            ndev = lib.evo_device_count()
        except backend.BackendError as e:
            raise RuntimeError(f"Compiler not found. Aborting. ({e})")
        if ndev <= 0:
            raise RuntimeError("Compiler not found. Aborting. (no CUDA device visible; the B200 backend has no CPU fallback)")
        self._device = (mpi_rank % ndev) if device is None else device
        self._device_problems: Dict[Tuple, backend.DeviceProblem] = {}
        self._average_generation_time = 0
        self._counter = 0
        self.last_outcome = None
        self.total_kernel_launches = 0
        # what generate_primitive_set needs (exastencils_FAS.py:76-93)
        try:
            import sympy
            from evostencils.grammar import multigrid as mg
            from evostencils.ir import base
            self.fields = [sympy.Symbol("u")]
            self.equations, self.operators = [], []
            for i in range(self.min_level, self.max_level + 1):
                self.equations.append(mg.EquationInfo("solEq", i, f"( Laplace@{i} * u@{i} ) == RHS_u@{i}"))
                self.operators.append(mg.OperatorInfo("RestrictionNode", i, None, base.Restriction))
                self.operators.append(mg.OperatorInfo("ProlongationNode", i, None, base.Prolongation))
                self.operators.append(mg.OperatorInfo("Laplace", i, None, base.Operator))
            size = 2 ** self.max_level
            self.finest_grid = [base.Grid((size,) * self.dimension, (1.0 / size,) * self.dimension, self.max_level)]
        except Exception:
            self.fields = [SimpleNamespace(name="u")]
            self.equations, self.operators = [], []
            size = 2 ** self.max_level
            self.finest_grid = [SimpleNamespace(size=(size,) * self.dimension, spacing=(1.0 / size,) * self.dimension,
                                                level=self.max_level)]
        self.coarsening_factor = [tuple([2] * self.dimension)]

    @property
    def uses_FAS(self):
        return True

    @property
    def problem(self):
        return self._problem

    def _dev(self, min_level, max_level):
        key = (min_level, max_level)
        if key not in self._device_problems:
            self._device_problems[key] = backend.DeviceProblem(self._problem.with_levels(min_level, max_level), self._device)
        return self._device_problems[key]

    @staticmethod
    def _find_cycle(args):
        expression = None
        for arg in args:
            if type(arg).__name__ == "Cycle":
                expression = arg
        return expression

    def lower(self, expression) -> ol.Program:
        from . import lowering_fas
        p = self._problem
        lo = lowering_fas.FASLowering(0, p.max_level, p.dim, p.settings.cgs_max_iters, p.settings.damping)
        # the coarsest level of the individual is whatever its tree reaches
        lo.update_rhs = {l: True for l in range(0, p.max_level + 1)}
        lo._traverse(expression)
        levels = [o.level for o in lo.ops] + [o.level - 1 for o in lo.ops if o.code in (
            ol.OP_FAS_RESTRICT_SOL, ol.OP_FAS_COARSE_RHS, ol.OP_PROLONG_ADD)]
        lo_level = min(levels)
        prog = ol.Program(dim=p.dim, n_fields=1, min_level=lo_level, max_level=p.max_level, ops=lo.ops,
                          operators={l: p.operator(l) for l in range(lo_level, p.max_level + 1)})
        prog.restrict_w, prog.prolong_w = p.restrict_weights(), p.prolong_weights()
        return prog

    def generate_and_evaluate(self, *args, **kwargs):
        infinity = 1e100
        expression = self._find_cycle(args)
        samples = kwargs.get("evaluation_samples", 1)
        start = time.time()
        try:
            prog = self.lower(expression)
            dev = self._dev(prog.min_level, prog.max_level)
            s = dev.problem.settings
            cyc = dev.build(prog)
            try:
                out = cyc.solve(s.tol, s.max_iters, samples=max(1, int(samples)))
            finally:
                cyc.close()
        except backend.BackendError as e:
            if e.infrastructure:                                # not a property of the individual
                raise
            return infinity, infinity, infinity
        except lowering.LoweringError:
            return infinity, infinity, infinity
        self._counter += 1
        self._average_generation_time += (time.time() - start - self._average_generation_time) / self._counter
        self.last_outcome = out
        self.total_kernel_launches += out.kernel_launches * max(1, int(samples))
        t, c, n = fitness.fas_fitness(out.residuals, out.time_ms, infinity)
        if self.cumulative_timer and n < infinity:
            t = t * (n + 1) / 2.0
        return t, c, n

    def generate_cycle_function(self, *args):
        expression = self._find_cycle(args)
        prog = self.lower(expression)
        return exaslang.program_to_exaslang(prog, ("Solution",), ("RHS",), self.max_level, self._cycle_name)

    def generate_storage(self, *args):
        return []

    def initialize_code_generation(self, *args):
        return None

    def close(self):
        for d in self._device_problems.values():
            d.close()
        self._device_problems.clear()


ProgramGeneratorFAS = B200ProgramGeneratorFAS
