#!/bin/bash
# Dev helper for one `gpurun --gpus N` call: the multi-GPU tests and the N-GPU bench line.  usage: tools/run_multi_gpu_checks.sh <tag> <N>
tag=${1:-x}; n=${2:-8}
out=gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $out/${tag}_gpu_tests_${n}gpu_multi.log 2>&1; tail -3 $out/${tag}_gpu_tests_${n}gpu_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 5 --warmup 3 > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_${n}gpu.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"], d["n_gpus"], d["config"]["gallery_total"], d["parity"], d["comm"], d["e2e"])
print({k:(v["value"],v["ms_per_step"],v["parity"]["bit_identical"],v["parity"].get("top100_identical")) for k,v in (d.get("sub") or {}).items()})
PY
grep -c "NCCL INFO" $out/${tag}_bench_${n}gpu.err; grep -m2 "nranks" $out/${tag}_bench_${n}gpu.err | cut -c1-200
