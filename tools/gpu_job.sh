python -m pytest tests/test_monoflex_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_tests4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_tests4.log
for c in monoflex km3d gac; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench4_$c.json 2> gpurun_out/r2_bench4_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench4_$c.json').read().strip().splitlines()[-1]);print('$c', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])"; done
python tools/error_budget.py > gpurun_out/r2_error_budget.txt 2>&1; cat gpurun_out/r2_error_budget.txt
for mb in 0 44; do VD3D_TC_L2MB=$mb ncu --set full --clock-control none -k regex:conv2d_tcph_kernel -s 49 -c 9 --csv --page raw --log-file gpurun_out/r2_ncu_wide_l2mb$mb.csv python bench.py --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_ncu_wide_l2mb$mb.log 2>&1; done
echo done
