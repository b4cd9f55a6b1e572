set -x
cd /root/repo
mkdir -p gpurun_out
for cfg in "0 3" "1 3" "3 3" "3 4" "2 3"; do
  set -- $cfg
  EFTB_RESUM_VARIANT=$1 EFTB_RESUM_MINB=$2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ab_v$1_m$2.json 2> gpurun_out/ab_v$1_m$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_v$1_m$2.json").read().strip().splitlines()[-1])
    print("VARIANT $1 MINB $2", d["ms_per_step"], d["stage_ms"]["resum"], d["logp_check"])
except Exception as e:
    print("VARIANT $1 MINB $2 failed", e)
PY
done
EFTB_RESUM_VARIANT=3 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v3.log 2>&1; tail -3 gpurun_out/pytest_gpu_v3.log
