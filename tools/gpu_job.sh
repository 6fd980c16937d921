timeout 300 python tools/exp_bottleneck.py 10 > gpurun_out/r2_exp_bottleneck3.log 2>&1; grep "default tile policy\|bn_tile = 64" gpurun_out/r2_exp_bottleneck3.log | cut -c1-150
run() { env "$@" timeout 300 python bench.py --config $CFG --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_b27.json 2> gpurun_out/r2_b27.err; python -c "
import json;d=json.load(open('gpurun_out/r2_b27.json'));print('$CFG $*',d['value'],d['e2e']['value'],d['clocks']['sm_mhz'])"; tail -1 gpurun_out/r2_b27.err | cut -c1-200; }
CFG=gac; run A=0; run VD3D_TC_SHORTK=0; run VD3D_TC_SHORTK_RES=0; run A=0; run VD3D_TC_SHORTK=0; run VD3D_TC_SHORTK_RES=0
CFG=monoflex; run A=0; run VD3D_TC_SHORTK_RES=0
CFG=stereo; run A=0; run VD3D_TC_SHORTK_RES=0
timeout 600 python -m pytest tests/test_mono3d_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/r2_tests27.log 2>&1; echo "pytest mono3d rc=$?"; tail -3 gpurun_out/r2_tests27.log | cut -c1-300
