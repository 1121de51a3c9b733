# round 2, call AC: streamed step with the next step's uploads released after the previous step's last E-step
# (TTLDA_UPLOAD_OVERLAP=1, default) against waiting for all queued work (=0), one GPU: streaming tests, alternating runs
set -x
O=gpurun_out/r02ac
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x -k "stream or host or e2e" > $O/pytest_stream.log 2>&1; echo "pytest rc=$?" >> $O/pytest_stream.log
tail -4 $O/pytest_stream.log | cut -c1-300
S="--steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check"
for V in 1 0 1 0 1 0; do
TTLDA_UPLOAD_OVERLAP=$V timeout 300 python bench.py $S > $O/bench.json 2> $O/bench.err
python - "$O/bench.json" $V <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("overlap", sys.argv[2], "e2e ms", round(j["e2e"]["ms_per_step"], 3), "int32", round(j["e2e"]["int32"]["ms_per_step"], 3), "value ms", round(j["ms_per_step"], 3))
PY
done
