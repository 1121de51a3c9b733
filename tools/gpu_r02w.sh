# round 2, call W: deep register pipelines in the E-step (TTLDA_SKEW=3: three row buffers, two sub-batches in flight, 10 one-warp
# (provenance of profiles/r02m_* / r02vw_*: the variants measured here were removed from the product afterwards — tools/experiments/estep_row_pipelines.cuh)
# CTAs per SM; =4: four buffers + skewed chain, 9 per SM) against the default walk (two buffers, 12 warps), same box
set -x
O=gpurun_out/r02w
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "deep_register" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sel.log; tail -6 $O/pytest_sel.log | cut -c1-300
S="--steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check --no-e2e"
timeout 300 python bench.py $S > $O/bench_base.json 2> $O/bench_base.err
TTLDA_SKEW=3 timeout 300 python bench.py $S > $O/bench_skew3.json 2> $O/bench_skew3.err
TTLDA_SKEW=4 timeout 300 python bench.py $S > $O/bench_skew4.json 2> $O/bench_skew4.err
timeout 300 python bench.py $S > $O/bench_base2.json 2> $O/bench_base2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02w/bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(j["ms_per_step"], 3), j.get("kernel_ms"), "perp", j["perplexity_trace"][-1])
    except Exception as e:
        print(f, "unreadable", e)
PY
