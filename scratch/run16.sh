#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests -x -q -m gpu -k "window_filters or cxx or reference or chain" 2>&1 | tail -15
python - <<'PY'
import json
d=json.load(open('gpurun_out/reference_suite_report.json'))
print({k:(v['passed'],v['run']) for k,v in d.items() if isinstance(v,dict) and 'run' in v})
PY
