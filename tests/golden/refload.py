"""Import the UNMODIFIED reference (read-only mount /root/reference) in this container, on CPU.

Only used by the golden-vector generator scripts in this directory (never at test/bench time:
/root/reference does not exist on the GPU box).  Follows the recipe in SURVEY.md section 8(c):
environmental shims only, no reference source edits.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("VISUALDET3D_REF", "/root/reference")


class EasyDict(dict):
    """Minimal stand-in for the uninstalled `easydict` package."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {})
        d.update(kw)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def copy(self):
        return EasyDict(dict.copy(self))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    """Returns the imported `visualDet3D` reference package (CPU-runnable)."""
    if "visualDet3D" in sys.modules and getattr(sys.modules["visualDet3D"], "_b200_ref", False):
        return sys.modules["visualDet3D"]
    import torch
    import torchvision

    os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")
    sys.dont_write_bytecode = True
    _stub("easydict", EasyDict=EasyDict)
    for n in ("skimage", "skimage.io", "skimage.measure", "matplotlib", "matplotlib.pyplot"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                _stub(n)
    # compiled extensions are import-time dependencies only (no GPU here): stub the two pybind modules
    notimpl = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("reference CUDA extension stub"))
    _stub("visualDet3D.networks.lib.ops.dcn.deform_conv_ext",
          deform_conv_forward=notimpl, deform_conv_backward_input=notimpl,
          deform_conv_backward_parameters=notimpl, modulated_deform_conv_forward=notimpl,
          modulated_deform_conv_backward=notimpl)
    _stub("visualDet3D.networks.lib.ops.iou3d.iou3d_cuda",
          boxes_iou_bev_gpu=notimpl, boxes_overlap_bev_gpu=notimpl, nms_normal_gpu=notimpl, nms_gpu=notimpl)
    # hard-coded .cuda() / cuda.synchronize() in the reference -> no-ops on a CPU-only host
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    sys.path.insert(0, REF_ROOT)
    import visualDet3D
    import visualDet3D.networks  # registers detectors
    # DCNv2 has no CPU path in the reference (deform_conv.py:174-175): torchvision stand-in, same mmcv lineage
    from visualDet3D.networks.lib.ops.dcn import deform_conv as _dc

    def _mdcn_cpu(x, off, m, w, b, s=1, p=0, d=1, g=1, dg=1):
        return torchvision.ops.deform_conv2d(x, off, w, b, stride=s, padding=p, dilation=d, mask=m)

    def _dcn_cpu(x, off, w, s=1, p=0, d=1, g=1, dg=1):
        return torchvision.ops.deform_conv2d(x, off, w, None, stride=s, padding=p, dilation=d)

    _dc.modulated_deform_conv = _mdcn_cpu
    _dc.deform_conv = _dcn_cpu
    visualDet3D._b200_ref = True
    return visualDet3D
