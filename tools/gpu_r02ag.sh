# round 2, call AG: final state — full GPU test suite, smoke(), the default bench line
set -x
O=gpurun_out/r02ag
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log | cut -c1-300
timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02ag/bench_default.json").read().strip().splitlines()[-1])
print("ms/step", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["ms_per_step"], j["e2e"]["value"], "int32", j["e2e"]["int32"]["ms_per_step"], "frac", j["roofline"]["frac"])
for k, v in j["extra_configs"].items():
    print(k, v["ms_per_step"], v["kernel_ms"])
PY
