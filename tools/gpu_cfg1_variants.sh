set -u
mkdir -p gpurun_out
run() { out=gpurun_out/v_$1.json
  env $2 timeout 300 python bench.py --workload cfg1 --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 2 > $out 2> ${out%.json}.err
  python -c "import json;d=json.load(open('$out'));print('$1', 'step %.4f ms apply %.4f ms pack %.4f ms frac %.3f same=%s'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['config']['pack_ms_per_step'],d['roofline']['frac'],d['config']['device_vs_host_path_identical']))" || tail -3 ${out%.json}.err; }
run default "A=1"
run variant2_sell "RG_APPLY_VARIANT=2"
run slices_w4 "RG_APPLY_VARIANT=4"
run w8 "RG_GROUP_WIDTH=8"
run slices_w8 "RG_APPLY_VARIANT=4 RG_GROUP_WIDTH=8"
