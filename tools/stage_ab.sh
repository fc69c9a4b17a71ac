#!/bin/bash
# A/B of the per-stage device times of bench.py under environment switches (one JSON line each).
# usage: tools/stage_ab.sh "VAR=1" "OTHER=1 X=2" ...   (an empty string = defaults)
for cfg in "$@"; do
  echo "== [$cfg]"
  env $cfg python bench.py --steps 12 --warmup 4 --no-cpu --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print({k: round(v, 4) for k, v in d['stage_ms'].items()}, 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'issue', d['host_issue_ms'])
    elif 'Error' in l or 'error' in l:
        print(l.strip())
"
done
