# usage: bash profiles/ncu_launches.sh <tag>  -- per-launch device times of one bench run (cold-cache, serialised: compare shares)
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config3 --no-batch --no-band"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_l_$TAG.log
