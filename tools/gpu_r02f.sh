# round 2, call F (8 GPUs): the N=8 bench line — cfg3 weak, cfg4 strong, cfg5 x 8 (sampling=True) — with the fused
# peer-memory M-step, and cfg4 strong again with the round-1 exchange (all-reduce + replicated M-step)
set -x
O=gpurun_out/r02f
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT $RUN 29541 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench rc=$?"
grep -i "nvls\|channels\|Connected" $O/bench_n8.err | head -12 > $O/nccl_info.txt
tail -c 1200 $O/bench_n8.err
TTLDA_MSTEP=replicated $RUN 29542 bench.py --gpus 8 --config cfg4 --strong --no-extra --no-e2e --no-parity --no-fit-check --steps 5 --warmup 3 > $O/bench_n8_cfg4_repl.json 2> $O/bench_n8_cfg4_repl.err
TTLDA_MSTEP=nccl $RUN 29543 bench.py --gpus 8 --config cfg4 --strong --no-extra --no-e2e --no-parity --no-fit-check --steps 5 --warmup 3 > $O/bench_n8_cfg4_nccl.json 2> $O/bench_n8_cfg4_nccl.err
tail -c 400 $O/bench_n8_cfg4_repl.err $O/bench_n8_cfg4_nccl.err
