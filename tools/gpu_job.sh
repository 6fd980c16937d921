python -m pytest tests/test_monoflex_gpu.py tests/test_zz_next_rows_gpu.py tests/test_dcn_iou3d_gpu.py tests/test_reference_seam.py tests/test_stereo3d_gpu.py -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/r2_tests2.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_tests2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_stereo.csv python bench.py --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_prof_stereo.log 2>&1
for mb in 0 28 44 64; do VD3D_TC_L2MB=$mb python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_l2mb_$mb.json 2> gpurun_out/r2_bench_l2mb_$mb.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_l2mb_$mb.json').read().strip().splitlines()[-1]);print('L2MB=$mb', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks'].get('power_w'))"; done
for c in gac monoflex km3d yolo3d; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_$c.json').read().strip().splitlines()[-1]);print('$c', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_monoflex.csv python bench.py --config monoflex --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_prof_monoflex.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_gac.csv python bench.py --config gac --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_prof_gac.log 2>&1
echo done
