# usage: bash profiles/ncu_kernels.sh <tag> <kernel-regex> [skip] [count]  -- full ncu capture of selected kernels
TAG=${1:-x}; RE=${2:-bsel_main}; SKIP=${3:-10}; CNT=${4:-3}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_$TAG.log
