timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "row_strip" > gpurun_out/r2_tests13.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_tests13.log | cut -c1-300
for v in 1 0 1 0; do VD3D_STEM_ROWS=$v timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench13_$v.json 2> gpurun_out/r2_bench13_$v.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench13_$v.json'));print('STEM_ROWS=$v',d['value'],d['e2e']['value'],d['clocks']['sm_mhz'])"; tail -2 gpurun_out/r2_bench13_$v.err; done
