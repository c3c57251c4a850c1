# parity tests of ONE study build, then the A/B timing of all of them:  VARIANT=E bash tools/studies/ab_test.sh
BMPC_LIB=$PWD/tools/studies/build/ab_${VARIANT}.so python -m pytest tests/test_gpu_linmpc.py tests/test_golden.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -3
ROUNDS=${ROUNDS:-2} bash tools/studies/ab.sh
