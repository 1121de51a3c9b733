# round 2, call U (2 GPUs): N > 1 sanity with the slice-walk mirror — 2-GPU parity under pytest, N = 2 bench line (weak cfg3 +
# strong cfg4 extras, e2e per rank through the streamed step)
set -x
O=gpurun_out/r02u
mkdir -p $O
timeout 900 python -m pytest tests/test_named_configs.py -m gpu -q -k "two_gpus" > $O/pytest_two_gpus.log 2>&1; echo "pytest rc=$?" >> $O/pytest_two_gpus.log
tail -4 $O/pytest_two_gpus.log | cut -c1-200
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 1200 $RUN 29551 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"
tail -c 500 $O/bench_n2.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02u/bench_n2.json").read().strip().splitlines()[-1])
print("N=2 ms/step", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["ms_per_step"], "parity", j.get("multigpu_parity"))
for k, v in (j.get("extra_configs") or {}).items():
    print(k, v.get("ms_per_step"), v.get("value"))
PY
