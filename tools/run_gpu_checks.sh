#!/bin/bash
# Dev helper for one gpurun call.  usage: tools/run_gpu_checks.sh <tag> [steps...]
# steps: tests bench hard harder ingest ncu_all launches sub
tag=${1:-x}; shift
out=gpurun_out
for step in "$@"; do
case $step in
tests) python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_gpu_tests.log;;
bench) python bench.py --no-sub > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?";;
sub) python bench.py > $out/${tag}_bench_full.json 2> $out/${tag}_bench_full.err; echo "bench full rc=$?";;
hard) python bench.py --no-sub --profile hard > $out/${tag}_bench_hard.json 2> $out/${tag}_bench_hard.err; echo "hard rc=$?";;
harder) python bench.py --no-sub --profile harder > $out/${tag}_bench_harder.json 2> $out/${tag}_bench_harder.err; echo "harder rc=$?"; tail -c 300 $out/${tag}_bench_harder.err;;
ingest) python tools/bench_ingest.py 2000 100000 > $out/${tag}_ingest.json 2> $out/${tag}_ingest.err; echo "ingest rc=$?"; tail -c 300 $out/${tag}_ingest.err; cat $out/${tag}_ingest.json;;
ncu_all) ncu --set full --clock-control none --import-source on -k regex:"tex_rowmax|minu_sim_kernel|minu_select_kernel|graph_minu_sparse|graph_tex_sparse" -s 5 -c 5 -o $out/prof_${tag}_all -f python bench.py --steps 1 --warmup 1 --no-sub --parity-sample 64 > $out/ncu_${tag}.log 2>&1; tail -2 $out/ncu_${tag}.log;;
launches) ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(tex_|minu_|graph_|fuse_|topk_|keys_to_hits|merge_hits|fill_empty)" -c 400 --csv --log-file $out/launches_${tag}.csv python bench.py --steps 2 --warmup 1 --no-sub --parity-sample 64 > $out/launches_${tag}.log 2>&1; tail -1 $out/launches_${tag}.log | cut -c1-200;;
sanitizer) compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_compnet.py tests/test_pq_encode.py -m gpu -x -q > $out/${tag}_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -4 $out/${tag}_sanitizer.log;;
esac
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/${tag}_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity"], d["exactness"])
        print(d["roofline"]["kernel_ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
