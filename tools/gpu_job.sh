set -x
python -m pytest tests/test_gpu_ops.py -q -x 2>&1 | tail -2
python -m pytest tests/test_gpu_query.py -q -x -k "query_parity" 2>&1 | tail -2
for tws in 0 1; do for c in 24 84 148 273 2960; do APSU_B200_NTT_TWS=$tws python tools/bench_ntt.py 16M-4096 $c | sed "s/^/TWS=$tws /"; done; done
for c in 84 148 296; do for tws in 0 1; do APSU_B200_NTT_TWS=$tws python tools/bench_ntt.py 1M-1024-cmp $c | sed "s/^/TWS=$tws /"; done; done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus 2 --workload 16M-1024 --db-log2 22 --steps 30 --warmup 5 --no-dag-split > gpurun_out/r2_split_none.json 2> gpurun_out/r2_split_none.err
$TR bench.py --gpus 2 --workload 16M-1024 --db-log2 22 --steps 30 --warmup 5 --dag-split > gpurun_out/r2_split_p2p.json 2> gpurun_out/r2_split_p2p.err
APSU_B200_NO_P2P=1 $TR bench.py --gpus 2 --workload 16M-1024 --db-log2 22 --steps 30 --warmup 5 --dag-split > gpurun_out/r2_split_nccl.json 2> gpurun_out/r2_split_nccl.err
tail -c 600 gpurun_out/r2_split_p2p.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_split_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j['ms_per_step'],3), 'e2e', round(j['e2e']['ms_per_step'],3), j['scopes_ms_rank0_last_step'], j['config']['parallelism'][-70:], j.get('parity_sample',{}).get('bit_exact_vs_oracle'), j.get('INVALID'))
    except Exception as e: print(f,'ERR',e)
PY
