#!/bin/bash
# e2e-only A/B: label|env|extra bench args
for spec in "$@"; do
  IFS='|' read -r label envs bargs <<< "$spec"
  out=$(env $envs timeout 900 python bench.py --no-cpu --no-others --steps 4 --warmup 1 --streams 8 --batch 8 $bargs 2>> gpurun_out/ab.err)
  python - "$label" "$out" <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2]); e = d["e2e"]
    print("%-24s e2e %8.1f fps  single %s  parse1 %.1f" % (sys.argv[1], e["value"], e.get("single_stream_fps"), d["run"]["host_parse_fps_1thread"]))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex, sys.argv[2][:200])
PY
done
