"""Drop-in module named `apriltag`.

The reference does `from apriltag import apriltag` (src/detection/tag_detector.py:7-11,
src/simulation/simulation_engine.py:24-26).  Put this package directory on sys.path (or
`sys.modules['apriltag'] = aprilslam_b200.apriltag`) and tag_detector.py runs unmodified on the B200.
Also exposes the pip-style `Detector()` used by scripts/verify_installation.py:45-51.
"""
try:
    from .detector import apriltag, records_to_dicts, Detector as _Detector
except ImportError:  # imported as a top-level module from sys.path
    from aprilslam_b200.detector import apriltag, records_to_dicts, Detector as _Detector


class Detector:
    """pip `apriltag.Detector()`-style facade (scripts/verify_installation.py:45-51): detect(gray) -> dicts."""

    def __init__(self, families="tag36h11", **kw):
        self._d = apriltag(families, **kw)

    def detect(self, gray):
        return list(self._d.detect(gray))


__all__ = ["apriltag", "Detector"]
