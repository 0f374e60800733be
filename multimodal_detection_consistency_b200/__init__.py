"""B200-native TVC scoring + retrieval hot path (libtvc.so + the reference-shaped host API)."""
from ._native import (Context, DetectorParams, Gallery, TvcError, default_params, load_library,  # noqa: F401
                      EXPORTED_SYMBOLS, SCORE_NAMES, SCORE_INDEX, NSCORES)

__version__ = "0.1.0"
