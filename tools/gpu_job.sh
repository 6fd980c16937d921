timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_tests_final3.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_tests_final3.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 4 > gpurun_out/r2_bench_final3.json 2> gpurun_out/r2_bench_final3.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_final3.json'));print('stereo',d['value'],d['e2e']['value'],d['e2e_f32']['value'],d['roofline']['frac'],d['clocks']['sm_mhz'],d['cpu_baseline']['value'])"; tail -2 gpurun_out/r2_bench_final3.err | cut -c1-300
for c in gac monoflex km3d yolo3d; do timeout 300 python bench.py --config $c --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench_final3_$c.json 2> gpurun_out/r2_bench_final3_$c.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_final3_$c.json'));print('$c',d['value'],d['e2e']['value'],d['ms_per_step'])"; done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_final3_ref.json 2> gpurun_out/r2_bench_final3_ref.err; tail -c 600 gpurun_out/r2_bench_final3_ref.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
