"""Op modules the reference builds under make.sh, re-exposed over libvd3d_b200 with the same function names and argument
meaning: `ops.dcn` (pybind module `deform_conv_ext`, R/lib/ops/dcn/src/deform_conv_ext.cpp:149-163 + the Python wrappers of
deform_conv.py) and `ops.iou3d` (`iou3d_cuda`, R/lib/ops/iou3d/src/iou3d.cpp:174-179 + iou3d.py)."""
from . import dcn, iou3d  # noqa: F401
