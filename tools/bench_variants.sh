#!/bin/bash
# throughput of bench.py under environment variants: tools/bench_variants.sh "VAR=1 VAR2=x" ...
for v in "$@"; do
  env $v python bench.py --no-cpu-baseline --steps ${STEPS:-20} 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$v', 'F=%d' % d['config']['in_flight_solves'], 'value %.4gM' % (d['value'] / 1e6), 'ms %.3f' % d['ms_per_step'], 'e2e %.4gM' % (d['e2e']['value'] / 1e6), 'serial %.4gM' % (d['config']['serial']['value'] / 1e6))"
done
