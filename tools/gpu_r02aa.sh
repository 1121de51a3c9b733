# round 2, call AA: full-size property tests (cfg3, cfg4) and the E-step launch granularity of the streamed step with the
# faster block mirror (TTLDA_ESTEP_GROUP = document blocks per E-step launch; default 2 for compact host formats)
set -x
O=gpurun_out/r02aa
mkdir -p $O
timeout 1200 python -m pytest tests/test_named_configs.py -m gpu -q -x -k "full_size" > $O/pytest_full_size.log 2>&1; echo "pytest rc=$?" >> $O/pytest_full_size.log
tail -15 $O/pytest_full_size.log | cut -c1-300
S="--steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check"
for G in 2 1 4 8 16 2; do
  TTLDA_ESTEP_GROUP=$G timeout 300 python bench.py $S > $O/bench_group$G.json 2> $O/bench_group$G.err
  python - "$O/bench_group$G.json" $G <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("group", sys.argv[2], "e2e ms", round(j["e2e"]["ms_per_step"], 3), "int32", round(j["e2e"]["int32"]["ms_per_step"], 3), "value ms", round(j["ms_per_step"], 3))
PY
done
