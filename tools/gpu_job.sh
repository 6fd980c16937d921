python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_tests_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; tail -c 700 gpurun_out/r2_bench_final.json
for c in gac monoflex km3d yolo3d; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_final_$c.json 2> gpurun_out/r2_bench_final_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_final_$c.json').read().strip().splitlines()[-1]);print('$c', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_prof_final.log 2>&1
ncu --set full --clock-control none -k regex:'conv2d_tc' -s 132 -c 44 --csv --page raw --log-file gpurun_out/r2_ncu_convs.csv python bench.py --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_ncu_convs.log 2>&1
ncu --set full --clock-control none -k regex:'psm_cosine|concat_conv3d|conv3d_2|sort_nms|decode_cand|pool_border|split_h16|dwconv|avgpool|image_to' -s 60 -c 20 --csv --page raw --log-file gpurun_out/r2_ncu_misc.csv python bench.py --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_ncu_misc.log 2>&1
ncu --set full --clock-control none -k regex:'deform_conv_fused|look_ground|post_opt' -s 0 -c 4 --csv --page raw --log-file gpurun_out/r2_ncu_gac.csv python bench.py --config gac --profile-mode --steps 1 --warmup 1 > gpurun_out/r2_ncu_gac.log 2>&1
ncu --set full --clock-control none -k regex:'deform_conv_fused' -s 48 -c 16 --csv --page raw --log-file gpurun_out/r2_ncu_dcn.csv python bench.py --config monoflex --profile-mode --steps 1 --warmup 3 > gpurun_out/r2_ncu_dcn.log 2>&1
echo done
