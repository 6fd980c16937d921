python -m pytest tests/test_ops_gpu.py tests/test_stereo3d_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "stem or range or stereo3d or Stereo or against or full_size or engines or pipeline or record or lo_companions or protocol or decode" > gpurun_out/r2_tests7.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_tests7.log
for v in "VD3D_X=0" "VD3D_STEM_POOL=0" "VD3D_TC_XMAJOR=0" "VD3D_X=1"; do env $v python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('stereo $v', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['power_w'])"; done
for c in monoflex gac; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench7_$c.json 2> gpurun_out/r2_bench7_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench7_$c.json').read().strip().splitlines()[-1]);print('$c', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])"; tail -2 gpurun_out/r2_bench7_$c.err; done
echo done
