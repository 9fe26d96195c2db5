#!/bin/bash
for w in 4 8 16; do
  DGADJ_HOST_WAVES=$w timeout 250 python bench.py --steps 2 --warmup 3 --no-cpu 2>/dev/null > /tmp/w$w.json
  python -c "
import json
d=json.loads(open('/tmp/w$w.json').read().strip().splitlines()[-1])
print('waves $w value %.4e e2e %.4e ms %.1f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
