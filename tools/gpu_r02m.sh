# round 2, call M: slice-walk mirror v2 (plain 16-bit counters + atomic redo, coalesced colptr scan, slice-staged placement),
# (provenance of profiles/r02m_* / r02vw_*: the variants measured here were removed from the product afterwards — tools/experiments/estep_row_pipelines.cuh)
# cp.async ring-fed E-step A/B at cfg3 / cfg4 / cfg5 shard
set -x
O=gpurun_out/r02m
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slice_walk or cp_async_ring" > $O/pytest_sel.log 2>&1
RC=$?; echo "pytest rc=$RC" >> $O/pytest_sel.log; tail -15 $O/pytest_sel.log | cut -c1-300
S="--steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check"
timeout 600 python bench.py $S > $O/bench_cfg3_base.json 2> $O/bench_cfg3_base.err
TTLDA_RING=1 timeout 600 python bench.py $S > $O/bench_cfg3_ring.json 2> $O/bench_cfg3_ring.err
TTLDA_CSC_WALK=0 timeout 600 python bench.py $S > $O/bench_cfg3_cubsort.json 2> $O/bench_cfg3_cubsort.err
timeout 600 python tools/diag_e2e.py > $O/diag_e2e_walk.txt 2>&1
X="--steps 3 --warmup 3 --no-cpu --no-extra --no-fit-check --no-e2e"
timeout 600 python bench.py --config cfg4 $X > $O/bench_cfg4_base.json 2> $O/bench_cfg4_base.err
TTLDA_RING=1 timeout 600 python bench.py --config cfg4 $X > $O/bench_cfg4_ring.json 2> $O/bench_cfg4_ring.err
timeout 600 python bench.py --config cfg5 --docs 1250000 $X > $O/bench_cfg5_base.json 2> $O/bench_cfg5_base.err
TTLDA_RING=1 timeout 600 python bench.py --config cfg5 --docs 1250000 $X > $O/bench_cfg5_ring.json 2> $O/bench_cfg5_ring.err
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(j["ms_per_step"], 3), "kernels", j.get("kernel_ms"), "e2e", (j.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log | cut -c1-300
