#!/bin/bash
# A/B replay throughput on one GPU box: each line of gpurun_out/ab.txt = label + bench.py JSON.
# usage: ab.sh "<label>|<env assignments>|<bench.py args>" ...
mkdir -p gpurun_out
: > gpurun_out/ab.txt
for spec in "$@"; do
  IFS='|' read -r label envs bargs <<< "$spec"
  echo "== $label" >> gpurun_out/ab.err
  out=$(env $envs timeout 600 python bench.py --no-cpu --no-e2e --no-others $bargs 2>> gpurun_out/ab.err)
  echo "$label $out" >> gpurun_out/ab.txt
  python - "$label" "$out" <<'PY'
import json, sys
label, out = sys.argv[1], sys.argv[2]
try:
    d = json.loads(out)
    st = d["roofline"]["all"]
    print("%-28s value %9.1f fps  step %8.2f ms  submit %7.2f ms | " % (label, d["value"], d["ms_per_step"], d["run"]["host_submit_ms_per_step"]) +
          " ".join("%s %.2f" % (k.replace("k_", ""), v["ms_per_clip_pass"]) for k, v in st.items()) +
          " | dom %s frac %.4f" % (d["roofline"]["kernel"], d["roofline"]["frac"]))
except Exception as e:
    print(label, "FAILED", e, out[:300])
PY
done
