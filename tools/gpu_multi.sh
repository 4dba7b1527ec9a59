#!/bin/bash
# multi-GPU bench: $1 = N, remaining args: env assignments (e.g. MB_BENCH_NCCL=1)
N=$1; shift
mkdir -p gpurun_out
env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/n$N.err | tail -1 > gpurun_out/bench_n$N.json
grep "cached iteration" gpurun_out/n$N.err | head -2; grep -i "error\|Traceback" gpurun_out/n$N.err | head -5
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print("N=$N $*", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "pose_err", d["check"]["final_pose_err_m"], "e2e pose err", d["e2e"].get("final_pose_err_m"))
PY
