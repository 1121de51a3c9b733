# round 2, call AE (8 GPUs): weak cfg3 bench line, every rank holding its own documents of ONE corpus (same topics), final streamed step (block mirror by slice walk, uploads released
# by the previous step's last E-step)
set -x
O=gpurun_out/r02ae
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
timeout 900 $RUN 29561 bench.py --gpus 8 --steps 10 --warmup 3 --no-extra --no-parity --no-fit-check --no-cpu > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench rc=$?"
tail -c 300 $O/bench_n8.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02ae/bench_n8.json").read().strip().splitlines()[-1])
print("N=8 ms/step", j["ms_per_step"], "value", j["value"], "e2e ms", j["e2e"]["ms_per_step"], "int32", j["e2e"]["int32"]["ms_per_step"], "e2e value", j["e2e"]["value"])
print(j.get("per_rank"))
PY
