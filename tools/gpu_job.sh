python - <<'PY'
import os, sys, json, time
os.environ["APSU_B200_BUILD_TIMING"]="1"
sys.path.insert(0,'.')
import bench, apsu_b200, torch
pj=bench.load_params_json("16M-4096")
params=apsu_b200.PSUParams.Load(json.dumps(pj))
for i in range(2):
    print("---- run", i, flush=True)
    r=bench.full_db_build_measure(pj, params, 0, 24)
    print({k:r[k] for k in ("ms","ms_first","bin_bundles")}, flush=True)
PY
