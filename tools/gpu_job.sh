timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "row_strip or stem" > gpurun_out/r2_tests15.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests15.log | cut -c1-300
timeout 200 python tools/exp_stem.py > gpurun_out/r2_exp_stem.log 2>&1; cat gpurun_out/r2_exp_stem.log | tail -8
for v in 0 1 0 1; do VD3D_PDL=$v timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench15_$v.json 2> gpurun_out/r2_bench15_$v.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench15_$v.json'));print('PDL=$v',d['value'],d['e2e']['value'],d['clocks']['sm_mhz'])"; tail -2 gpurun_out/r2_bench15_$v.err; done
VD3D_PDL=1 timeout 600 python -m pytest tests/test_stereo3d_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/r2_tests15b.log 2>&1; echo "pytest PDL rc=$?"; tail -3 gpurun_out/r2_tests15b.log | cut -c1-300
