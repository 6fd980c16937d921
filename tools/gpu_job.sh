timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_last.json 2> gpurun_out/r2_bench_last.err; echo "rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_last.json'));print(d['value'],d['warmup'],d['steps'],d['config']['warmup_steps_run'],d['e2e']['value'],d['gpu_launches'])"; tail -2 gpurun_out/r2_bench_last.err | cut -c1-200
