"""The drop-in boundary EXECUTED against the real reference package.

CPU part (here and on any box that has the reference tree or the oracle/_ref copy): `plugin.install_into_reference()`, construction
through the reference's registry from an EasyDict config for all five detectors, strict `load_state_dict` of state dicts produced by
the reference's own modules.  GPU part (-m gpu, tests/workers/seam_gpu.py): the `sys.modules` substitution of `deform_conv_ext` /
`iou3d_cuda` under the reference's UNMODIFIED deform_conv.py / iou3d.py, the reference's unmodified `test_stereo_detection` /
`test_mono_detection` pipelines driving the B200 classes, and the unmodified reference detector run on the same GPU.
Workers run in their own process: importing the reference changes torch / sys.modules globally."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refload  # noqa: E402

needs_ref = pytest.mark.skipif(not refload.available(), reason="no reference package (neither /root/reference nor oracle/_ref/visualDet3D)")


def run_worker(name, timeout=900):
    env = dict(os.environ)
    env.pop("VD3D_CONV_ENGINE", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "workers", name)], capture_output=True, text=True, timeout=timeout, env=env)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("SEAM_JSON ")]
    assert r.returncode == 0 and lines, f"worker {name} failed (rc {r.returncode}):\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    return json.loads(lines[-1][len("SEAM_JSON "):])


@needs_ref
def test_install_into_reference_and_strict_state_dicts():
    out = run_worker("seam_cpu.py")
    assert out["installed"] == sorted(["Stereo3D", "Yolo3D", "GroundAwareYolo3D", "MonoFlex", "KM3D"])
    assert out["pipelines_kept"]
    assert set(out["reference_detectors_left"]) >= {"RetinaNet", "MonoDepth"}            # untouched: out of scope, still the reference's
    for name, rec in out["detectors"].items():
        assert rec["class_module"].startswith("visualdet3d_b200."), (name, rec)
        assert rec["is_nn_module"] and rec["keys_equal"] and rec["shapes_equal"], (name, rec)
        assert rec["missing"] == [] and rec["unexpected"] == [] and rec["values_loaded"], (name, rec)
        assert rec["n_params"] == rec["n_params_reference"] > 1_000_000, (name, rec)
        assert rec["cpu_forward"] != "ran", (name, rec)                                  # the B200 classes have no CPU path
    print({k: v["n_params"] for k, v in out["detectors"].items()})


@needs_ref
@pytest.mark.gpu
def test_reference_modules_and_pipelines_over_b200_ops():
    out = run_worker("seam_gpu.py", timeout=1500)
    print(json.dumps(out, indent=1))
    # unmodified ModulatedDeformConvPack / DeformConvPack of the reference, running on visualdet3d_b200.ops.dcn vs on its own extension
    assert out["dcn_v2_rel_err"] < 1e-4 and out["dcn_v1_rel_err"] < 1e-4
    assert out["dcn_v2_grad_rel_err"] < 1e-4 and out["dcn_v1_grad_rel_err"] < 1e-4      # the reference's autograd Functions over our backward entries
    assert out["iou3d_max_err"] < 1e-5 and out["iou3d_nms_equal"]
    # the reference's unmodified test pipelines driving the B200 detectors == the committed reference fixtures
    for k in ("stereo", "mono"):
        assert out[k]["count"] == out[k]["fixture_count"] > 0 and out[k]["names_ok"]
        assert out[k]["max_score_diff"] < 1e-3 and out[k]["max_box_diff"] < 1e-3
    # the UNMODIFIED reference detector on the same GPU (its own CUDA ops, fp32 without TF32) vs the B200 class
    assert out["ref_gpu"]["count"] == out["ref_gpu"]["b200_count"] > 0
    assert out["ref_gpu"]["max_score_diff"] < 1e-3 and out["ref_gpu"]["max_box_diff"] < 1e-3
