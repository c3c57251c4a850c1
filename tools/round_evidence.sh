# Round evidence on ONE B200: the GPU suite, the default bench line (all configs), the reference arm, the ncu launch list and
# one full capture of the dominant kernel.  Outputs under gpurun_out/ (copied into profiles/ by hand after reading them).
R=${R:-r02}
python -m pytest tests -m gpu -q > gpurun_out/${R}_tests.log 2>&1; tail -2 gpurun_out/${R}_tests.log
python bench.py 2> gpurun_out/${R}_bench_n1.err | grep "^{" > gpurun_out/${R}_bench_n1.json
python bench.py --impl reference 2> gpurun_out/${R}_bench_ref.err | grep "^{" > gpurun_out/${R}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 20 --warmup 3 --no-configs > gpurun_out/${R}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_warp -s 30 -c 1 -f -o gpurun_out/warp_${R} python bench.py --steps 20 --warmup 3 --no-configs > gpurun_out/${R}_ncu_full.log 2>&1
ls -la gpurun_out/warp_${R}.ncu-rep
python -c "
import json; d=json.load(open('gpurun_out/${R}_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'], {k:(v.get('value'), v.get('e2e',{}).get('value')) for k,v in d.get('configs',{}).items()})"
