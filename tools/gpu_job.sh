set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err; tail -c 400 gpurun_out/r2_final_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err
python tools/cpu_full_query.py --threads all,1 > gpurun_out/cpu_full_query_r02.json 2> gpurun_out/cpu_full_query.err
export APSU_B200_NO_GRAPH=1
python tools/profile_query.py > gpurun_out/plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_final.csv python tools/profile_query.py > gpurun_out/ncu1.log 2>&1
python tools/profile_query.py --only-idx 0 > gpurun_out/plain2.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_one_bundle_index.csv python tools/profile_query.py --only-idx 0 > gpurun_out/ncu2.log 2>&1
python tools/profile_query.py > gpurun_out/plain3.log 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_ks_ -c 6 -o gpurun_out/prof_r02_ks python tools/profile_query.py > gpurun_out/ncu3.log 2>&1
unset APSU_B200_NO_GRAPH
python - <<'PY'
import json
for f in ('gpurun_out/r2_final_n1.json','gpurun_out/r2_final_ref.json'):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j['ms_per_step'],3), 'e2e', j['e2e'].get('ms_per_step'), j.get('e2e_seeded',{}).get('ms_per_step'), j.get('scopes_ms_rank0_last_step'), 'K1', j.get('roofline',{}).get('frac'), j.get('parity_sample',{}).get('bit_exact_vs_oracle'), j.get('results_sha256_matches_n1_record'), j.get('INVALID'), j.get('db_build_full'))
    except Exception as e: print(f,'ERR',e)
print(open('gpurun_out/cpu_full_query_r02.json').read()[:1500])
PY
