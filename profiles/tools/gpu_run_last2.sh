#!/bin/bash
# full -m gpu suite on the default build, A/B of the late PDL wait (default build vs -DTDVC_PDL_LATE_WAIT=0), then the final
# stage1 line (inference sweep + CPU baseline) on the faster of the two
mkdir -p gpurun_out
(timeout 200 python -m pytest tests -m gpu -q --tb=short --maxfail=30 -p no:cacheprovider > gpurun_out/last2_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/last2_tests.log)
tail -3 gpurun_out/last2_tests.log
NOLATE=td-vc-gan_b200/tdvc/libtdvc_b200_nolate.so
ab() {
  env "$2" timeout 120 python bench.py --steps 10 --warmup 3 --no-inference --no-cpu-baseline > gpurun_out/last2_$1.json 2> gpurun_out/last2_$1.err
  python -c "import json; print(json.load(open('gpurun_out/last2_$1.json'))['ms_per_step'])" 2>/dev/null || echo 999
}
m1=$(ab late TDVC_AB=late); m0=$(ab nolate TDVC_LIB=$NOLATE)
echo "late wait: $m1 ms   wait at the top: $m0 ms"
if python -c "import sys; sys.exit(0 if float('$m1') <= float('$m0') else 1)"; then echo "final line on the default build"; else echo "final line on the nolate build"; export TDVC_LIB=$NOLATE; fi
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/final3_stage1.json 2> gpurun_out/final3_stage1.err; echo "stage1 exit $?"
python -c "
import json; d=json.load(open('gpurun_out/final3_stage1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['inference']['best'])"
