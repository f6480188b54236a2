B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config3 --no-batch --no-band"
P='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], d["ms_per_step"], ["%.3f"%x for x in d["roofline"]["stages_ms"].values()], d["config"]["last_step_shift_m"])'
$B | python -c "$P" list
DCR_FRONT_INLINE=1 $B | python -c "$P" inline
$B | python -c "$P" list
DCR_FRONT_INLINE=1 $B | python -c "$P" inline
DCR_FRONT_INLINE=1 python -m pytest tests/test_gpu_parity.py -q -x -k "iteration or front or loop" 2>&1 | tail -3
