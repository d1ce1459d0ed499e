set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2d_n8.json 2> gpurun_out/r2d_n8.err; tail -c 300 gpurun_out/r2d_n8.err
$TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 --no-dag-split --no-parity > gpurun_out/r2d_n8_nosplit.json 2> gpurun_out/r2d_n8_nosplit.err
$TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2d_n4.json 2> gpurun_out/r2d_n4.err
$TR --nproc-per-node 2 --master-port 29525 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2d_n2.json 2> gpurun_out/r2d_n2.err
$TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --workload 256M-4096 --db-log2 28 --steps 10 --warmup 3 > gpurun_out/r2d_256M_n8.json 2> gpurun_out/r2d_256M_n8.err; tail -c 300 gpurun_out/r2d_256M_n8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['ms_per_step'],3), 'shared', j.get('e2e_shared_query',{}).get('ms_per_step'), j['config']['parallelism'][-60:], j.get('parity_sample',{}).get('bit_exact_vs_oracle'), j.get('results_sha256_matches_n1_record'), j.get('INVALID'))
        print('    per-rank ms', [round(r['ms_per_step'],3) for r in j.get('per_rank',[])])
    except Exception as e: print(f,'ERR',e)
PY
