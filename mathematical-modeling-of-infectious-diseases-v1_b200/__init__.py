"""sepaihrd_b200 -- B200-native batched SEPAIHRD Dopri5 + Poisson-likelihood evaluator.

The directory name (``mathematical-modeling-of-infectious-diseases-v1_b200``) is not a valid Python
identifier; ``__graft_entry__.load_package()`` registers it as the module ``sepaihrd_b200``.

Contents: the problem description and reference-format readers (pure numpy, importable without a
GPU), and -- in :mod:`.capi` / :mod:`.evaluator` -- the ctypes binding of the C-ABI CUDA library
``csrc/libsepaihrd_b200.so``.  There is no CPU fallback: creating an evaluator without the built
library or without a CUDA device raises.
"""
from .problem import (Problem, SlotLayout, CProblem, load_default_problem, default_problem_path,
                      CLAMP, REFLECT, LOWEST, NUM_COMPARTMENTS, TRAJ_FULL, TRAJ_OBSERVED,
                      ST_OK, ST_S_OVERFLOW, ST_STEP_FAILURE, ST_NONFINITE, ST_INVALID_PARAM)
from . import config

__all__ = ["Problem", "SlotLayout", "CProblem", "load_default_problem", "default_problem_path", "config",
           "CLAMP", "REFLECT", "LOWEST", "NUM_COMPARTMENTS", "TRAJ_FULL", "TRAJ_OBSERVED",
           "ST_OK", "ST_S_OVERFLOW", "ST_STEP_FAILURE", "ST_NONFINITE", "ST_INVALID_PARAM"]
__version__ = "0.1.0"
