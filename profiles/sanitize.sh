# usage: bash profiles/sanitize.sh <tag>   -- compute-sanitizer (memcheck, racecheck, initcheck, synccheck) over small
# invocations of every kernel family of the hot path (profiles/sanitize_driver.py); logs under gpurun_out/sanitize_<tool>_<tag>.log
TAG=${1:-x}
mkdir -p gpurun_out
python profiles/sanitize_driver.py > gpurun_out/sanitize_plain_$TAG.log 2>&1 || { tail -5 gpurun_out/sanitize_plain_$TAG.log; exit 1; }
for tool in memcheck racecheck initcheck synccheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no"
  [ $tool = initcheck ] && extra="--track-unused-memory no"
  timeout 1500 compute-sanitizer --tool $tool $extra --print-limit 20 --log-file gpurun_out/sanitize_${tool}_$TAG.log \
      python profiles/sanitize_driver.py > gpurun_out/sanitize_${tool}_$TAG.out 2>&1
  echo "== $tool: exit $?"; tail -3 gpurun_out/sanitize_${tool}_$TAG.log; tail -2 gpurun_out/sanitize_${tool}_$TAG.out
done
