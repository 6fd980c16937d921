"""Detectors registered under the reference's names in `visualdet3d_b200.plugin.DETECTOR_DICT`."""
from .stereo3d import Stereo3D, build_synthetic_stereo3d  # noqa: F401
