timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 500 --warmup 20 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("default lib", round(d["ms_per_step"]*1e3,1), "us  loss", d["config"]["loss_check"])'
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --no-e2e > /dev/null 2>&1
grep -E "prep_kernel|final_kernel" gpurun_out/launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | head
