timeout 600 python -m pytest tests/test_monoflex_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/r2_tests17.log 2>&1; echo "pytest monoflex rc=$?"; tail -6 gpurun_out/r2_tests17.log | cut -c1-300
for v in 1 0 1 0; do VD3D_ROWCONV=$v timeout 300 python bench.py --config monoflex --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench17_$v.json 2> gpurun_out/r2_bench17_$v.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench17_$v.json'));print('ROWCONV=$v',d['value'],d['e2e']['value'],d['clocks']['sm_mhz'],d['gpu_launches'])"; tail -2 gpurun_out/r2_bench17_$v.err; done
timeout 300 python bench.py --config km3d --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench17_km3d.json 2> gpurun_out/r2_bench17_km3d.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench17_km3d.json'));print('km3d',d['value'],d['e2e']['value'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches17_monoflex.csv python bench.py --config monoflex --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_prof17.log 2>&1; tail -1 gpurun_out/r2_prof17.log
