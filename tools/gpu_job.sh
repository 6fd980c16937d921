python -m pytest tests/test_zz_next_rows_gpu.py tests/test_stereo3d_gpu.py tests/test_mono3d_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "graph or pipeline or streamed" > gpurun_out/r2_tests11.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_tests11.log
python tools/exp_chunk.py > gpurun_out/r2_exp_chunk.log 2>&1; cat gpurun_out/r2_exp_chunk.log | cut -c1-400
for c in gac monoflex yolo3d; do python bench.py --config $c --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench11_$c.json 2> gpurun_out/r2_bench11_$c.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench11_$c.json'));print('$c',d['value'],d['e2e']['value'],d['gpu_launches'])"; tail -2 gpurun_out/r2_bench11_$c.err; done
python bench.py --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err;  python -c "
import json;d=json.load(open('gpurun_out/r2_bench11.json'));print('stereo',d['value'],d['e2e']['value'],d['gpu_launches'])"; tail -2 gpurun_out/r2_bench11.err
