timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 4 > gpurun_out/r2_bench_n2c.json 2> gpurun_out/r2_bench_n2c.err; echo "bench rc=$?"; python -c "
import json
for l in open('gpurun_out/r2_bench_n2c.json'):
    if l.startswith('{'):
        d=json.loads(l); print('n2',d['value'],d['e2e']['value'],d['gather_verified'])
    else: print('EXTRA STDOUT LINE:',l[:80])"; grep -c "NCCL version" gpurun_out/r2_bench_n2c.err
