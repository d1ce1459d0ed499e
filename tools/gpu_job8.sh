set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_n8.json 2> gpurun_out/r2_n8.err; tail -c 300 gpurun_out/r2_n8.err
$TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 --no-dag-split --no-parity > gpurun_out/r2_n8_nosplit.json 2> gpurun_out/r2_n8_nosplit.err
$TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_n4.json 2> gpurun_out/r2_n4.err
python bench.py --workload 256M-4096 --db-log2 28 --steps 5 --warmup 3 --no-cpu-baseline --no-db-build --write-digest > gpurun_out/r2_256M_n1.json 2> gpurun_out/r2_256M_n1.err; tail -c 300 gpurun_out/r2_256M_n1.err
$TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --workload 256M-4096 --db-log2 28 --steps 10 --warmup 3 > gpurun_out/r2_256M_n8.json 2> gpurun_out/r2_256M_n8.err; tail -c 300 gpurun_out/r2_256M_n8.err
cp profiles/results_sha256_* gpurun_out/
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_n*.json')+glob.glob('gpurun_out/r2_256M*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['ms_per_step'],3), 'shared', j.get('e2e_shared_query',{}).get('ms_per_step'), j['scopes_ms_rank0_last_step'], 'K1', round(j['roofline']['frac'],3), j['config']['parallelism'][-60:], j.get('parity_sample',{}).get('bit_exact_vs_oracle'), j.get('results_sha256_matches_n1_record'), j.get('INVALID'))
    except Exception as e: print(f,'ERR',e)
PY
