# usage: bash profiles/ab_variants.sh <tag> [variant.so ...]  -- A/B timing of library builds under demcoreg_b200/csrc/variants/
# (DCR_LIB selects the build; default build first).  Prints ms/step and the per-stage CUDA-event times.
TAG=${1:-x}; shift
mkdir -p gpurun_out
for v in default "$@"; do
  if [ "$v" = default ]; then unset DCR_LIB; else export DCR_LIB=$PWD/demcoreg_b200/csrc/variants/$v.so; fi
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-config3 --no-batch --no-band > gpurun_out/ab_${TAG}_$v.json 2> gpurun_out/ab_${TAG}_$v.err
  python - gpurun_out/ab_${TAG}_$v.json $v <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    st = list(d['roofline']['stages_ms'].values())
    print("%-12s %.4f ms/step  stages %s  shift %s fb %s" % (sys.argv[2], d['ms_per_step'], " ".join("%.3f" % x for x in st),
          d['config']['last_step_shift_m'], d['config']['engines']['fast_fallback']))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
done 2>&1 | tee gpurun_out/ab_$TAG.txt
