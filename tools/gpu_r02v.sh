# round 2, call V: E-step occupancy A/B at cfg3 — 1-warp CTAs with a register cap (13 / 14 / 15 warps per SM at 152 / 144 /
# (provenance of profiles/r02m_* / r02vw_*: the variants measured here were removed from the product afterwards — tools/experiments/estep_row_pipelines.cuh)
# 136 registers, small spills) against the 3 x 4-warp baseline (12 warps, 166 registers), same box
set -x
O=gpurun_out/r02v
mkdir -p $O
S="--steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check --no-e2e"
for V in base w12x1 w13 w14 w15 w16 base2; do
  L=""; case $V in base|base2) ;; *) L="tools/lib/libttlda_$V.so";; esac
  TTLDA_LIB=$L timeout 300 python bench.py $S > $O/bench_$V.json 2> $O/bench_$V.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02v/bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(j["ms_per_step"], 3), j.get("kernel_ms"), "perp", j["perplexity_trace"][-1])
    except Exception as e:
        print(f, "unreadable", e)
PY
