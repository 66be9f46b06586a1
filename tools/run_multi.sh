#!/bin/bash
# tools/run_multi.sh N [tag] — the multi-GPU measurements of one box (run under `gpurun --gpus N`):
#   BASELINE config 5 (one 64 GB input block-sharded over N GPUs), the host<->device copy ceiling of the box at N,
#   and the default weak-scaling bench line at N.  Outputs land in gpurun_out/<tag>_*_n<N>.json.
N=${1:-2}
TAG=${2:-r02}
PORT=$((29600 + N))
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT"
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  python -m pytest tests -m gpu -q -k "two_contexts" 2>&1 | tail -3 > gpurun_out/${TAG}_two_contexts.log
  cat gpurun_out/${TAG}_two_contexts.log
fi
$RUN tools/pcie_ceiling.py > gpurun_out/${TAG}_pcie_n${N}.json 2> gpurun_out/${TAG}_pcie_n${N}.err
cat gpurun_out/${TAG}_pcie_n${N}.json
$RUN bench.py --gpus $N --workload cfg5 --steps 3 --warmup 1 > gpurun_out/${TAG}_cfg5_n${N}.json 2> gpurun_out/${TAG}_cfg5_n${N}.err
tail -c 400 gpurun_out/${TAG}_cfg5_n${N}.err
cat gpurun_out/${TAG}_cfg5_n${N}.json
$RUN bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
tail -c 300 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_n${N}.json").read().strip().splitlines()[-1])
    print("bench n=${N}:", d["value"], d["e2e"], d["decompress"]["value"], d["decompress"]["e2e"], d["decompress"].get("reference_written"))
except Exception as e:
    print("bench parse failed", e)
PY
