python -m pytest tests/test_dcn_iou3d_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "fused" > gpurun_out/r2_tests8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests8.log
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$1', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks'].get('power_w'))"; }
(cd build/r1 && run "round-1 tree (a)")
run "current (a)"
(cd build/r1 && run "round-1 tree (b)")
run "current (b)"
VD3D_PLANES=0 run "current PLANES=0"
VD3D_STEM_POOL=0 run "current STEM_POOL=0"
VD3D_TC_XMAJOR=0 run "current XMAJOR=0"
VD3D_TC_L2MB=0 run "current L2MB=0"
VD3D_PLANES=0 VD3D_STEM_POOL=0 VD3D_TC_XMAJOR=0 VD3D_TC_L2MB=0 run "current all-off"
python tools/exp_dcn.py 64 96 320 8 > gpurun_out/r2_exp_dcn64b.log 2>&1; cat gpurun_out/r2_exp_dcn64b.log
python tools/exp_dcn.py 128 48 160 8 > gpurun_out/r2_exp_dcn128b.log 2>&1; head -4 gpurun_out/r2_exp_dcn128b.log
python bench.py --config monoflex --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench8_monoflex.json 2> gpurun_out/r2_bench8_monoflex.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench8_monoflex.json').read().strip().splitlines()[-1]);print('monoflex', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])"; tail -2 gpurun_out/r2_bench8_monoflex.err
echo done
