timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python tools/replay.py bench_1440p 1 --stages 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --streams 32 --e2e-streams 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ALL', d['value'], d['ms_per_step'], d['roofline']['per_stage_ms_per_clip_pass'])"
