set -x
python -m pytest tests/test_gpu_query.py -q -k "delivered or facade or mask" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --write-digest > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 600 gpurun_out/r2_bench_n1.err
APSU_B200_NO_FUSE=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-db-build --no-parity > gpurun_out/r2_bench_n1_nofuse.json 2>&1
for w in "256K-512 18" "1M-1024-cmp 20" "1M-4096-com 20"; do set -- $w; python bench.py --workload $1 --db-log2 $2 --steps 50 --warmup 5 --no-cpu-baseline --no-db-build > gpurun_out/r2_bench_$1.json 2>gpurun_out/r2_bench_$1.err; APSU_B200_NO_FUSE=1 python bench.py --workload $1 --db-log2 $2 --steps 50 --warmup 5 --no-cpu-baseline --no-db-build --no-parity > gpurun_out/r2_bench_$1_nofuse.json 2>/dev/null; done
cp profiles/results_sha256_* gpurun_out/ 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j['ms_per_step'],3), 'e2e', round(j['e2e']['ms_per_step'],3), j['scopes_ms_rank0_last_step'], 'K1', round(j['roofline']['frac'],3), j.get('parity_sample',{}).get('bit_exact_vs_oracle'), j.get('INVALID'))
    except Exception as e: print(f,'ERR',e)
PY
