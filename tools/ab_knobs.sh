#!/bin/bash
# A/B runs of the tuning knobs (environment variables read by libeftb200) on the bench workload; prints
# ms/step, and the antidiag / ap / resum stage times.  Usage on the GPU box: bash tools/ab_knobs.sh "EFTB_AD_NB=1" "EFTB_AD_NB=15" ...
for kv in "$@"; do
  echo "== $kv"
  env $kv python bench.py --no-cpu --no-producer --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print(round(d['ms_per_step'],3), {k: s[k] for k in ('antidiag','spectral','resum','ap','likelihood')})"
done
