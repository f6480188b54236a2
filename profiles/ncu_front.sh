# usage: bash profiles/ncu_front.sh <tag>   (full ncu capture of the front-end kernels: fast interior-tile kernel + listed tiles)
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config3 --no-batch --no-band"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:front_ -s 6 -c 2 -o gpurun_out/prof_front_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
ls -la gpurun_out | tail -5
