#!/bin/bash
# tools/try_variant.sh "<-D flags>" <tag> — rebuild the library on the GPU box with extra compile flags, run the
# kernel-only bench, restore the default build.  Used for A/B runs of launch geometry.
FLAGS="$1"; TAG="$2"
cd fastqpacker_b200/csrc && touch *.cu && make -j16 -s EXTRA="$FLAGS" > /dev/null 2>&1 && cd ../..
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}.json").read().strip().splitlines()[-1])
print("${TAG} [${FLAGS}]:", round(d["value"], 1), "GB/s compress", round(d["decompress"]["value"], 1), "GB/s decompress", {k: v for k, v in d["stages_ms_per_step"].items()})
PY
cd fastqpacker_b200/csrc && touch *.cu && make -j16 -s > /dev/null 2>&1 && cd ../..
