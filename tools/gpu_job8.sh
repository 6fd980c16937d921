python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 4 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2_bench_n8.json; tail -3 gpurun_out/r2_bench_n8.err
python bench.py --gpus 1 --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench_n8_n1.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_n8_n1.json'));print('n1 on the same box',d['value'],d['e2e']['value'])"
