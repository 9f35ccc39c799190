timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
for w in nyu hires config1; do python bench.py --workload $w --steps 200 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["config"]["workload"][:40], round(d["ms_per_step"]*1e3,1), "us", round(d["value"]/1e9,2), "Gpx/s frac", round(d["roofline"]["frac"],3))'; done
