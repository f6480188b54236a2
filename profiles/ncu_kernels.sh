# usage: bash profiles/ncu_kernels.sh <tag> <kernel-regex> [skip] [count]   (full ncu capture of the matching kernels)
TAG=${1:-x}
RE=${2:-fb_passB_kernel}
SKIP=${3:-7}
CNT=${4:-7}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config3 --no-batch --no-band"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o gpurun_out/prof_k_$TAG $CMD > gpurun_out/ncu_k_$TAG.log 2>&1
ls -la gpurun_out | tail -3
