set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_n1.json 2> gpurun_out/r2c_n1.err; tail -c 300 gpurun_out/r2c_n1.err
for w in "256K-512 18" "1M-1024-cmp 20" "1M-4096-com 20"; do set -- $w; python bench.py --workload $1 --db-log2 $2 --steps 50 --warmup 5 --no-cpu-baseline --no-db-build --write-digest > gpurun_out/r2c_$1.json 2>gpurun_out/r2c_$1.err; done
cp profiles/results_sha256_* gpurun_out/
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j['ms_per_step'],3), 'e2e', round(j['e2e']['ms_per_step'],3), j.get('e2e_seeded',{}).get('ms_per_step'), j['scopes_ms_rank0_last_step'], 'K1', round(j['roofline']['frac'],3), j.get('parity_sample',{}).get('bit_exact_vs_oracle'), j.get('results_sha256_matches_n1_record'), j.get('INVALID'), j.get('roofline_int',{}).get('frac'), j.get('db_build_full',{}).get('ms'))
    except Exception as e: print(f,'ERR',e)
PY
