"""pyglm-b200: B200-native engine for pyglm's population-GLM ll/gradient and Gibbs delta-ll hot path."""
from .engine import (Dataset, EngineError, load_library, NLIN_EXP, NLIN_SOFTPLUS,
                     PATH_AUTO, PATH_FP64, PATH_TC, X_F32, X_F64, X_PLANES)

__all__ = ["Dataset", "EngineError", "load_library", "NLIN_EXP", "NLIN_SOFTPLUS",
           "PATH_AUTO", "PATH_FP64", "PATH_TC", "X_F32", "X_F64", "X_PLANES"]
