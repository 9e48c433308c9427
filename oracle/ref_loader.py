"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference modules by file path.

Used in the build container (where /root/reference exists) by
`tests/golden/make_golden.py` to generate golden vectors and by
`tests/test_oracle_vs_reference.py` to validate the restatement in
`oracle/semgate_oracle.py`.  Nothing under the product package imports this
file, and nothing that runs on the GPU box needs /root/reference.

The package import `scripts.semantic_gating` fails here (matplotlib is absent,
`scripts/semantic_gating/__init__.py:28`), but `place_recognition.py:20-28` and
`loop_closure_gate.py:11-13` import only numpy at module scope, so they load
stand-alone.
"""
from __future__ import annotations

import importlib.util
import os
import sys

REFERENCE_ROOT = os.environ.get("SEMGATE_REFERENCE_ROOT", "/root/reference")
_SG = os.path.join(REFERENCE_ROOT, "scripts", "semantic_gating")


def available() -> bool:
    return os.path.isfile(os.path.join(_SG, "place_recognition.py"))


def _load(name: str, fname: str):
    key = f"_semgate_ref_{name}"
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(_SG, fname))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod


def place_recognition():
    return _load("pr", "place_recognition.py")


def loop_closure_gate():
    return _load("lcg", "loop_closure_gate.py")


class _Identity:
    """Mix-in making `extract_descriptor` the identity, so that `query()` and
    `add_image()` can be driven with ready-made descriptors (the reference demo
    bypasses extraction the same way, `place_recognition.py:1015-1020`)."""

    def extract_descriptor(self, image):
        return image


def identity_vpr(descriptor_dim: int):
    PR = place_recognition()
    cls = type("IdentityVPR", (_Identity, PR.BasePlaceRecognition), {})
    return cls(descriptor_dim=descriptor_dim, device="cpu")


def semantic_place_recognition(descriptor_dim: int, similarity_threshold=0.5, min_time_gap=10.0):
    """A reference `SemanticPlaceRecognition` whose extractor is the identity.
    The constructor would instantiate MixVPR (and try to import torch models);
    `__new__` + manual attributes mirrors `place_recognition.py:826-828`."""
    PR = place_recognition()
    spr = PR.SemanticPlaceRecognition.__new__(PR.SemanticPlaceRecognition)
    spr.similarity_threshold = similarity_threshold
    spr.min_time_gap = min_time_gap
    spr.vpr = identity_vpr(descriptor_dim)
    return spr


def run_find_loop_closures(desc, ts, floors, similarity_threshold=0.5, min_time_gap=10.0,
                           k=10, enable_floor_gating=True):
    """Run the reference `find_loop_closures` verbatim.  `floors` may contain
    None.  Returns the list of reference `PlaceMatch` objects."""
    PR = place_recognition()
    spr = semantic_place_recognition(desc.shape[1] if len(desc) else 0,
                                     similarity_threshold, min_time_gap)
    for i in range(len(desc)):
        f = floors[i]
        spr.vpr.descriptors.append(PR.PlaceDescriptor(
            timestamp=float(ts[i]), descriptor=desc[i],
            floor_label=None if f is None else int(f)))
    return spr, spr.find_loop_closures(enable_floor_gating=enable_floor_gating, k=k)
