"""semgate — B200-native gated loop-closure candidate retrieval.

Drop-in for the retrieval path of the reference's `scripts/semantic_gating` package
(names re-exported there at `__init__.py:33-41,23-27`).  Everything on the path runs in
libsemgate's sm_100a kernels; importing this package does not need a GPU, using it does.
"""
from .place_recognition import (  # noqa: F401
    PlaceMatch, PlaceDescriptor, MatchArrays, BasePlaceRecognition, SemanticPlaceRecognition,
    MixVPR, SALAD, AnyLoc, CricaVPR,
)
from .loop_closure_gate import LoopClosureCandidate, SemanticLoopClosureGate  # noqa: F401

__version__ = "0.1.0"
