# Quick GPU check of a step_warp change (one call, ≈ 1 min): seed accuracy, the parity tests that run on the warp kernel, timing.
./tools/studies/micro/rcp_acc
python -m pytest tests/test_gpu_linmpc.py tests/test_golden.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -3
python tools/studies/nscale.py 4096 2>&1 | grep -v Warn | head -12
