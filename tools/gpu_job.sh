set -x
python -m pytest tests/test_gpu_query.py -q -x -k "fused or query_parity" 2>&1 | tail -2
python -m pytest tests/test_gpu_dbbuild.py -q -x 2>&1 | tail -2
for v in "0 0" "1 148" "1 296" "1 600" "1 100000"; do set -- $v; echo "FUSE=$1 MAX=$2"; APSU_B200_FUSE=$1 APSU_B200_FUSE_MAX=$2 python tools/profile_query.py --only-idx 0 --warmup 3 | tail -1; done
for v in "0 0" "1 148" "1 296" "1 100000"; do set -- $v; echo "256K FUSE=$1 MAX=$2"; APSU_B200_FUSE=$1 APSU_B200_FUSE_MAX=$2 python bench.py --workload 256K-512 --db-log2 18 --steps 50 --warmup 5 --no-cpu-baseline --no-db-build --no-parity | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['ms_per_step'], j['scopes_ms_rank0_last_step'])"; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['ms_per_step'], j['db_build_full'])"
