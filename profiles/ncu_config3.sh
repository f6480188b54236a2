# usage: bash profiles/ncu_config3.sh <tag>  -- launch list + full capture of the NCC / SAD kernels (config 3)
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python profiles/run_config3.py 2"
$CMD > gpurun_out/c3_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c3_launches_$TAG.csv $CMD > gpurun_out/c3_ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sad_pass|ncc_tile|sad_sample" -s 4 -c 4 -o gpurun_out/prof_c3_$TAG $CMD > gpurun_out/c3_ncu_k_$TAG.log 2>&1
ls -la gpurun_out | tail -4
