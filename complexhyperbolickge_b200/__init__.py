"""B200-native scoring hot path of ComplexHyperbolicKGE (FFTRotH / FFTRefH / FFTAttH)."""
from .models import CHYP_MODELS, FFTAttH, FFTRefH, FFTRotH, KGModel  # noqa: F401

__all__ = ["FFTRotH", "FFTRefH", "FFTAttH", "KGModel", "CHYP_MODELS"]
