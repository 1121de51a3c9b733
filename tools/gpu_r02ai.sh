# round 2, call AI (2 GPUs): the driver's N = 2 command on the final state (default bench line: weak cfg3 + parity + cfg4 strong)
set -x
O=gpurun_out/r02ai
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 900 $RUN 29571 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02ai/bench_n2.json").read().strip().splitlines()[-1])
print("N=2 ms/step", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["ms_per_step"], j["e2e"]["value"])
print("parity ok:", {k: v.get("ok") for k, v in j["multigpu_parity"].items() if isinstance(v, dict)})
print(j.get("per_rank"))
for k, v in (j.get("extra_configs") or {}).items():
    print(k, v.get("ms_per_step"), v.get("value"))
PY
