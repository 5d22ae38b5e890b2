#!/bin/bash
# 2-GPU data-parallel check: overlapped bucketed all-reduce (captured in the graph) against the flat all-reduce between
# graph segments.  Prints both JSON lines into gpurun_out/.
mkdir -p gpurun_out
for ov in 1 0; do
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 --overlap $ov > gpurun_out/r2_dp2_overlap$ov.json 2> gpurun_out/r2_dp2_overlap$ov.err
  echo "overlap=$ov exit $?"; tail -c 400 gpurun_out/r2_dp2_overlap$ov.err
done
python - <<'PY'
import json
for ov in (1, 0):
    try:
        d = json.load(open(f"gpurun_out/r2_dp2_overlap{ov}.json"))
        print("overlap", ov, d["ms_per_step"], d["value"], d["losses"])
    except Exception as e:
        print("overlap", ov, "ERR", e)
PY
