python -m pytest tests/test_gpu_multi.py -q -s 2>&1 | grep -E "info|passed|failed|Error|error" | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2e_n2.json 2> gpurun_out/r2e_n2.err
python -c "
import json
j=json.loads(open('gpurun_out/r2e_n2.json').read().strip().splitlines()[-1])
print(round(j['ms_per_step'],3), j['e2e']['ms_per_step'], j['e2e_root_scatter']['ms_per_step'], j.get('parity_sample',{}).get('bit_exact_vs_oracle'), j.get('results_sha256_matches_n1_record'), j.get('INVALID'))
"
