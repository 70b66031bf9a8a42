#!/bin/bash
# Dev helper for one gpurun call: GPU tests, the benchmark (i.i.d. and hard profile), ingest at 100,000 directory
# entries and an ncu capture of the selection kernel.  usage: tools/run_gpu_checks.sh <tag>
tag=${1:-x}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_gpu_tests.log
python bench.py --no-sub > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --no-sub --profile hard > $out/${tag}_bench_hard.json 2> $out/${tag}_bench_hard.err; echo "hard rc=$?"; tail -c 400 $out/${tag}_bench_hard.err
python tools/bench_ingest.py 2000 100000 > $out/${tag}_ingest.json 2> $out/${tag}_ingest.err; echo "ingest rc=$?"; tail -c 300 $out/${tag}_ingest.err; cat $out/${tag}_ingest.json
python - <<PY
import json
for f in ("$out/${tag}_bench.json","$out/${tag}_bench_hard.json"):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity"], d["exactness"])
        print(d["roofline"]["kernel_ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
ncu --set full --clock-control none --import-source on -k regex:"minu_select_kernel" -s 1 -c 1 -o $out/prof_${tag}_select -f python bench.py --steps 1 --warmup 1 --no-sub --parity-sample 64 > $out/ncu_${tag}.log 2>&1; tail -2 $out/ncu_${tag}.log
