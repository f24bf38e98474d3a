# device-resident step time of one or more workloads (no CPU leg, no rollout, no C5): quick A/B during kernel work
cd $GRAFT_REPO_ROOT
for wl in "$@"; do
  for rep in 1 2; do
    python bench.py --workload $wl --steps 3000 --warmup 20 --no-cpu --no-rollout --no-c5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$wl', round(d['ms_per_step']*1e3,2), 'us/step  frac', round(d['roofline']['frac'],4))"
  done
done
