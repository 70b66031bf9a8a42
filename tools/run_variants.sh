#!/bin/bash
# Dev helper for A/B runs in one gpurun call: every msu-latentafis_b200/lib/variants/<name>.so is put in place of the
# library in turn and benchmarked (bench.py --no-sub); results go to gpurun_out/var_<tag>_<name>.json.
tag=${1:-v}
lib=msu-latentafis_b200/lib
for so in $lib/variants/*.so; do
  name=$(basename $so .so)
  cp $so $lib/liblatentafis_b200.so
  python bench.py --no-sub --parity-sample 256 > gpurun_out/var_${tag}_${name}.json 2> gpurun_out/var_${tag}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/var_${tag}_${name}.json").read().strip().split("\n")[-1])
    print("${name}", round(d["ms_per_step"],3), d["parity"]["bit_identical"], {k:round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("${name}", "unreadable", e)
PY
done
