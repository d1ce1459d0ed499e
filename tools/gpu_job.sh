set -x
python -m pytest tests/test_gpu_multi.py tests/test_gpu_dbbuild.py tests/test_gpu_prng.py -q -s 2>&1 | grep -v "^$" | tail -25
APSU_B200_NO_P2P=1 python -m pytest tests/test_gpu_multi.py -q -s -k two_gpu 2>&1 | grep -E "info|passed|failed" | tail -8
