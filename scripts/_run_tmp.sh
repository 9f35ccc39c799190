for cfg in "0 0" "16 150" "16 250" "16 400" "8 250" "32 250" "32 400" "16 600"; do
  set -- $cfg
  export SNDE_FINE_ROWS=$1 SNDE_FINE_PERMILLE=$2
  python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-e2e 2>gpurun_out/err.txt | tail -1 > gpurun_out/o.json
  python -c 'import sys,json,os; d=json.load(open("gpurun_out/o.json")); print("fine", os.environ["SNDE_FINE_ROWS"], os.environ["SNDE_FINE_PERMILLE"], round(d["ms_per_step"]*1e3,1), "us")' || tail -3 gpurun_out/err.txt
done
