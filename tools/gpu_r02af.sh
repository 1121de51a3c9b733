# round 2, call AF: E-step launch granularity of the streamed step now that the uploads arrive under the previous step's
# statistics pass (TTLDA_ESTEP_GROUP = document blocks per E-step launch)
set -x
O=gpurun_out/r02af
mkdir -p $O
S="--steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check"
for G in 2 4 8 16 2; do
  TTLDA_ESTEP_GROUP=$G timeout 300 python bench.py $S > $O/bench_group$G.json 2> $O/bench_group$G.err
  python - "$O/bench_group$G.json" $G <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("group", sys.argv[2], "e2e ms", round(j["e2e"]["ms_per_step"], 3), "int32", round(j["e2e"]["int32"]["ms_per_step"], 3), "value ms", round(j["ms_per_step"], 3))
PY
done
