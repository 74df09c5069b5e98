# A/B of the recurrence's cluster budget inside the pipelined bench loops (PLAS_REC_MAX_CLUSTERS overrides the descriptor's budget)
for wl in c2 c4; do
for mc in 0 7 0 7; do
  if [ "$mc" != "0" ]; then export PLAS_REC_MAX_CLUSTERS=$mc; else unset PLAS_REC_MAX_CLUSTERS; fi
  python bench.py --workload $wl --steps ${STEPS:-20} --warmup 5 --no-cpu-baseline --no-sub-records 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl max_clusters override $mc', 'value', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'rec', round(d['stages']['rec']['ms_per_step'],3), d['stages']['rec'].get('rows_per_group'))"
done; done
