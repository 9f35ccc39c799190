timeout 900 python -m pytest tests -m gpu -q -k "frames or lean_instance_equals" 2>&1 | tail -2
timeout 600 python bench.py --steps 300 --warmup 30 --no-cpu-baseline 2>gpurun_out/bench.err | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("step", round(d["ms_per_step"]*1e3,1), "us; e2e", d["e2e"])'
