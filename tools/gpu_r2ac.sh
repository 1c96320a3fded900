#!/bin/bash
one() {
  timeout 300 python bench.py --workload c2 --no-cpu --no-compare "$@" 2>/dev/null > /tmp/x.json
  python - "$*" <<'PY'
import json,sys
d=json.loads([l for l in open('/tmp/x.json') if l.startswith('{')][-1]); print(sys.argv[1], round(d['value']), d['config']['slots_per_launch'], d['config']['launch_groups_per_step'])
PY
}
one
one --slots 300
one --slots 600
one --slots 600 --group 100
one --slots 600 --group 128
one --slots 600 --group 75
one --slots 384 --group 128
one --slots 256 --group 86
one
