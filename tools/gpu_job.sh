python -m pytest tests/test_dcn_iou3d_gpu.py tests/test_monoflex_gpu.py tests/test_mono3d_gpu.py -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/r2_tests3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests3.log
for s in layer1 layer2 layer3; do python tools/exp_conv.py $s 12 > gpurun_out/r2_exp_$s.log 2>&1; done; cat gpurun_out/r2_exp_layer1.log
for c in monoflex km3d; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3_$c.json 2> gpurun_out/r2_bench3_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench3_$c.json').read().strip().splitlines()[-1]);print('$c', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])"; done
VD3D_DCN_FUSED=0 python bench.py --config monoflex --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('monoflex unfused', round(d['value'],1), round(d['ms_per_step'],3))"
for mc in 160 256; do VD3D_PLANES_MAXC=$mc python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('stereo PLANES_MAXC=$mc', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_monoflex_fused.csv python bench.py --config monoflex --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_prof_monoflex_fused.log 2>&1
echo done
