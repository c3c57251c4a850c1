source <(sed -n '/^run()/,/^}/p' tools/studies/gather_diag.sh)
export NG=8
run off8 BMPC_FUSED_GATHER=off
run side8 BMPC_PULL_MODE=side
