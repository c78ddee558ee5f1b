"""dagsfm_b200 -- B200-native hot path of AIBluefisher/DAGSfM.

Python here is a thin host-side mirror of the reference's matcher / verifier /
bundle-adjuster interfaces over the C ABI (include/dagsfm_b200.h); all numeric
work runs in the hand-written sm_100a CUDA library.
"""
from ._lib import B2Error, MatchOptions, lib  # noqa: F401
from .matching import (SiftMatchGPU, SiftMatchingOptions, match_guided_sift_features_gpu,  # noqa: F401
                       match_sift_features_gpu)

from .verification import (Camera, TwoViewGeometryVerifier, TwoViewOptions,  # noqa: F401
                           TwoViewResult)
from .bundle_adjustment import BundleAdjuster, BundleAdjustmentOptions  # noqa: F401

__version__ = "0.1"
from .ba_config import BundleAdjustmentConfig, Reconstruction, pack_problem, unpack_problem  # noqa: F401,E402
from .retrieval import QueryOptions, VisualIndex, VocabSimilarityGraph, Vocabulary, make_vocabulary  # noqa: F401,E402
