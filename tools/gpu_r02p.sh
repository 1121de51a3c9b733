# round 2, call P: slice-walk mirror v3.1 (next-document prefetch in the walk, batched base staging in the placement)
set -x
O=gpurun_out/r02p
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slice_walk" > $O/pytest_sel.log 2>&1
RC=$?; echo "pytest rc=$RC" >> $O/pytest_sel.log; tail -15 $O/pytest_sel.log | cut -c1-300
python tools/diag_csc.py > $O/diag_csc_walk.txt 2>&1
TTLDA_CSC_WALK=0 python tools/diag_csc.py > $O/diag_csc_cubsort.txt 2>&1
cat $O/diag_csc_walk.txt $O/diag_csc_cubsort.txt | tail -4
DIAG_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:csc_' -s 5 -c 12 --csv --log-file $O/launches_csc.csv python tools/diag_csc.py > $O/ncu_csc.log 2>&1
S="--steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check"
timeout 600 python bench.py $S > $O/bench_cfg3_walk.json 2> $O/bench_cfg3_walk.err
timeout 600 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu --no-extra --no-fit-check > $O/bench_cfg4_walk.json 2> $O/bench_cfg4_walk.err
TTLDA_CSC_WALK=0 timeout 600 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu --no-extra --no-fit-check > $O/bench_cfg4_cubsort.json 2> $O/bench_cfg4_cubsort.err
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(j["ms_per_step"], 3), "e2e", (j.get("e2e") or {}).get("ms_per_step"), (j.get("e2e") or {}).get("int32", {}).get("ms_per_step"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
timeout 600 python tools/diag_e2e.py > $O/diag_e2e_walk.txt 2>&1; tail -7 $O/diag_e2e_walk.txt
