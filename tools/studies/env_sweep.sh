# Environment-only tuning sweep of the warp kernel on ONE box (nscale.py, N = 4096): slow-first threshold (BMPC_LONG), and the
# switches BMPC_WARP_OCC / BMPC_L2_PREFETCH / BMPC_STATIC_FIRST (pass the settings to try as arguments).
run() { echo "$1: $(env $1 python tools/studies/nscale.py 4096 2>/dev/null | head -1 | cut -c1-60)"; }
for s in ${@:-BMPC_X=default BMPC_LONG=14 BMPC_LONG=16 BMPC_LONG=18 BMPC_LONG=21 BMPC_LONG=99 BMPC_X=default}; do run $s; done
