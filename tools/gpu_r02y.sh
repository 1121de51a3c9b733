# round 2, call Y: final validation of HEAD — full GPU test suite, smoke(), a short default-shaped bench run — and the outputs of
# the round-1 gather microbenchmarks (tools/membench2: arithmetic enabled level by level; membench7: occupancy 1..3 CTAs/SM)
set -x
O=gpurun_out/r02y
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -3 $O/smoke.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check > $O/bench_cfg3_short.json 2> $O/bench_cfg3_short.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02y/bench_cfg3_short.json").read().strip().splitlines()[-1])
print("ms/step", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], "frac", j["roofline"]["frac"], j["kernel_ms"])
PY
timeout 600 tools/membench2 > $O/membench2.txt 2>&1; tail -20 $O/membench2.txt
timeout 600 tools/membench7 > $O/membench7.txt 2>&1; tail -20 $O/membench7.txt
