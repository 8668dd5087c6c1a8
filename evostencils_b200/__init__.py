"""B200-native multigrid fitness-evaluation backend for EvoStencils (evaluate path only).

Drop-in for ``evostencils.code_generation.exastencils.ProgramGenerator`` (reference:
exastencils.py:39): :class:`evostencils_b200.program_generator.B200ProgramGenerator`.
"""
__version__ = "0.1.0"
