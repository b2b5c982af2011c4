"""B200-native SASRec / CAST training-and-evaluation hot path (drop-in for the reference's models/sasrec.py and
models/cast_*.py behind the same model interface).  The arithmetic is hand-written sm_100a CUDA reached through the
C ABI in include/cast_b200.h; there is no CPU fallback — importing model classes works anywhere, constructing one
requires the built extension and a CUDA device."""
from ._lib import CastError, LIB_PATH, load_library  # noqa: F401
from .engine import MODELS, Engine, model_plan, sinusoid_table  # noqa: F401
from .models import (CAST1, CAST2, CAST3, CAST4, CAST5, CAST6, CAST7, CAST8, CAST9, SASRec,  # noqa: F401
                     build_model)

__all__ = ["SASRec", "CAST1", "CAST2", "CAST3", "CAST4", "CAST5", "CAST6", "CAST7", "CAST8", "CAST9", "MODELS",
           "build_model", "Engine", "load_library", "CastError"]
