python -m pytest tests/test_ops_gpu.py tests/test_stereo3d_gpu.py tests/test_monoflex_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_tests6.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_tests6.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench6.json').read().strip().splitlines()[-1]);print('stereo', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'], d['clocks']['sm_mhz'], d['roofline']['frac'])"; tail -2 gpurun_out/r2_bench6.err
for v in "VD3D_TC_XMAJOR=0" "VD3D_STEM_POOL=0" "VD3D_TC_XMAJOR=0 VD3D_STEM_POOL=0"; do env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('stereo $v', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"; done
python tools/exp_conv.py head 6 > gpurun_out/r2_exp_head2.log 2>&1; head -12 gpurun_out/r2_exp_head2.log
python tools/exp_conv.py layer1 8 > gpurun_out/r2_exp_layer1b.log 2>&1; head -4 gpurun_out/r2_exp_layer1b.log
python tools/exp_dcn.py 64 96 320 8 > gpurun_out/r2_exp_dcn64.log 2>&1; cat gpurun_out/r2_exp_dcn64.log
python tools/exp_dcn.py 128 48 160 8 > gpurun_out/r2_exp_dcn128.log 2>&1; cat gpurun_out/r2_exp_dcn128.log
for c in monoflex gac; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench6_$c.json 2> gpurun_out/r2_bench6_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench6_$c.json').read().strip().splitlines()[-1]);print('$c', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_stereo6.csv python bench.py --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_prof_stereo6.log 2>&1
echo done
