# Study: where the fused gather's per-period cost comes from (2 GPUs).  BMPC_PULL_MODE none = publish only, main = pull on
# the step's stream, side = pull on a side stream; BMPC_STATIC_FIRST / BMPC_PULL_LOCALBUF isolate the work-queue and the
# symmetric-memory store.
run() { # name, env...
name=$1; shift
env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NG:-2} --steps 50 --warmup 5 --no-configs 2>gpurun_out/gd_$name.err | grep "^{" > gpurun_out/gd_$name.json
grep -m1 "tail_ms" gpurun_out/gd_$name.err
python -c "
import json,sys; d=json.load(open('gpurun_out/gd_$name.json')); print('$name', d['value'], d['ms_per_step'], d['gather']['equals_ncclAllGather'])"
}
run off BMPC_FUSED_GATHER=off
run none_dyn BMPC_PULL_MODE=none
run none_static BMPC_PULL_MODE=none BMPC_STATIC_FIRST=1
run none_static_local BMPC_PULL_MODE=none BMPC_STATIC_FIRST=1 BMPC_PULL_LOCALBUF=1
run main BMPC_PULL_MODE=main
run side BMPC_PULL_MODE=side
run side_static BMPC_PULL_MODE=side BMPC_STATIC_FIRST=1
