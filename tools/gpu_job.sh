for v in default 512 300 0; do
  if [ $v = default ]; then unset VD3D_TC_PHALO_MAXC; else export VD3D_TC_PHALO_MAXC=$v; fi
  python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench12_$v.json 2> gpurun_out/r2_bench12_$v.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench12_$v.json'));print('PHALO_MAXC=$v',d['value'],d['e2e']['value'],d['clocks']['sm_mhz'])"; tail -2 gpurun_out/r2_bench12_$v.err
done
unset VD3D_TC_PHALO_MAXC
python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench12_again.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2_bench12_again.json'));print('default again',d['value'],d['e2e']['value'],d['clocks']['sm_mhz'])"
