python -m pytest tests/test_dcn_iou3d_gpu.py tests/test_monoflex_gpu.py tests/test_mono3d_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_tests5.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests5.log
for c in monoflex km3d; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench5_$c.json 2> gpurun_out/r2_bench5_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench5_$c.json').read().strip().splitlines()[-1]);print('$c', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])"; tail -2 gpurun_out/r2_bench5_$c.err; done
VD3D_DCN_STAGED=0 python bench.py --config monoflex --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('monoflex global-gather', round(d['value'],1), round(d['ms_per_step'],3))"
python tools/exp_conv.py head 8 > gpurun_out/r2_exp_head.log 2>&1; cat gpurun_out/r2_exp_head.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_monoflex_staged.csv python bench.py --config monoflex --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_prof_monoflex_staged.log 2>&1
ncu --set full --clock-control none -k regex:deform_conv_fused -s 48 -c 3 --csv --page raw --log-file gpurun_out/r2_ncu_dcn_fused.csv python bench.py --config monoflex --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_ncu_dcn_fused.log 2>&1
echo done
