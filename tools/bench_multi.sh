# usage: NG=8 bash tools/bench_multi.sh   -- the driver's multi-GPU command line, one JSON line into gpurun_out/
NG=${NG:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 100 --warmup 5 2> gpurun_out/r02_bench_${NG}gpu.err | grep "^{" > gpurun_out/r02_bench_${NG}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${NG}gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['gather']['equals_ncclAllGather'], d['e2e']['value'], d['clocks'])"
