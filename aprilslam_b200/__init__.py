"""aprilslam_b200 -- B200-native AprilTag detection + per-tag pose behind AprilSLAM's detector boundary.

Nothing here imports the CUDA library eagerly; `aprilslam_b200._lib.load()` does, and raises if
libaprilgpu.so is missing (there is no CPU fallback).
"""
__version__ = "0.1.0"
