"""Worker of tests/test_reference_seam.py (own process: importing the reference patches torch globally).  CPU only.

Executes the drop-in boundary against the REAL reference package: `plugin.install_into_reference()` puts the B200 detector classes into
the reference's own registries; `DETECTOR_DICT[cfg.detector.name](cfg.detector)` -- the exact expression of scripts/eval.py:37 /
train.py:87 -- then builds OUR class from the reference's own EasyDict config, and a state_dict produced by the REFERENCE's module of
the same name loads into it with strict key equality.  Prints one JSON line."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402
import refload  # noqa: E402


def main():
    refload.load_reference()
    from visualDet3D.networks.utils import registry as ref_registry
    from visualdet3d_b200 import plugin, synth
    import visualdet3d_b200.detectors as D  # noqa: F401  (registers the B200 detectors in plugin.DETECTOR_DICT)
    from visualdet3d_b200.detectors.centernet import km3d_cfg, monoflex_cfg
    names = ["Stereo3D", "Yolo3D", "GroundAwareYolo3D", "MonoFlex", "KM3D"]
    ref_classes = {n: ref_registry.DETECTOR_DICT[n] for n in names}
    ref_pipelines = dict(ref_registry.PIPELINE_DICT.module_dict)
    tmp = tempfile.mkdtemp()
    cfgs = {}
    pm, ps = synth.synth_priors(16, 3, ["Car", "Pedestrian"])
    d = os.path.join(tmp, "s"); os.makedirs(d)
    synth.write_priors(d, pm, ps, ["Car", "Pedestrian"])
    cfgs["Stereo3D"] = synth.stereo3d_cfg(d, ["Car", "Pedestrian"])
    pm, ps = synth.synth_priors(16, 2, ["Car"])
    d = os.path.join(tmp, "m"); os.makedirs(d)
    synth.write_priors(d, pm, ps, ["Car"])
    cfgs["Yolo3D"] = synth.mono3d_cfg(d, "Yolo3D", ["Car"], None)
    cfgs["GroundAwareYolo3D"] = synth.mono3d_cfg(d, "GroundAwareYolo3D", ["Car"], None)
    cfgs["MonoFlex"], cfgs["KM3D"] = monoflex_cfg(), km3d_cfg()
    # the reference's own modules first (their state_dicts are what a checkpoint of the reference holds)
    ref_sd = {}
    for n in names:
        torch.manual_seed(0)
        m = ref_classes[n](refload.to_edict(cfgs[n]))
        ref_sd[n] = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ref = plugin.install_into_reference()
    out = {"installed": sorted(k for k in ref.DETECTOR_DICT.module_dict if ref.DETECTOR_DICT[k].__module__.startswith("visualdet3d_b200")),
           "reference_detectors_left": sorted(k for k in ref.DETECTOR_DICT.module_dict if not ref.DETECTOR_DICT[k].__module__.startswith("visualdet3d_b200")),
           "pipelines_kept": sorted(ref_pipelines) == sorted(ref.PIPELINE_DICT.module_dict), "detectors": {}}
    for n in names:
        cfg = refload.to_edict(cfgs[n])                     # an EasyDict, like cfg.detector of the reference's config files
        det = ref.DETECTOR_DICT[cfg.name](cfg)              # scripts/eval.py:37
        rec = {"class_module": type(det).__module__, "is_nn_module": isinstance(det, torch.nn.Module)}
        ours = det.state_dict()
        rec["keys_equal"] = list(ours.keys()) == list(ref_sd[n].keys())
        rec["shapes_equal"] = all(tuple(ours[k].shape) == tuple(v.shape) for k, v in ref_sd[n].items() if k in ours)
        res = det.load_state_dict(ref_sd[n], strict=True)   # scripts/eval.py:42 uses strict=False; strict proves the key contract
        rec["missing"], rec["unexpected"] = list(res.missing_keys), list(res.unexpected_keys)
        rec["n_params"] = int(sum(p.numel() for p in det.parameters()))
        rec["n_params_reference"] = int(sum(v.numel() for k, v in ref_sd[n].items() if k in dict(det.named_parameters())))
        rec["values_loaded"] = all(torch.equal(det.state_dict()[k], v) for k, v in ref_sd[n].items())
        det.eval(); det.train(); str(det)                   # the calls scripts/train.py / eval.py make on the object
        try:
            det([torch.zeros(1, 3, 32, 32), torch.zeros(1, 3, 4)])
            rec["cpu_forward"] = "ran"
        except Exception as e:                              # no CPU fallback: must fail loudly
            rec["cpu_forward"] = type(e).__name__
        out["detectors"][n] = rec
    print("SEAM_JSON " + json.dumps(out))


if __name__ == "__main__":
    main()
