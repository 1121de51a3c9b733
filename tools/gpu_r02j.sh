# round 2, call J (8 GPUs): N=8 bench line with the multicast M-step; cfg4 strong A/B: multicast / peer loads / nccl / replicated
set -x
O=gpurun_out/r02j
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
$RUN 29541 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench rc=$?"
tail -c 600 $O/bench_n8.err
C="--gpus 8 --config cfg4 --strong --no-extra --no-e2e --no-parity --no-fit-check --steps 10 --warmup 3"
$RUN 29542 bench.py $C > $O/bench_n8_cfg4_mc.json 2> $O/bench_n8_cfg4_mc.err
TTLDA_MULTICAST=0 $RUN 29543 bench.py $C > $O/bench_n8_cfg4_peer.json 2> $O/bench_n8_cfg4_peer.err
TTLDA_MSTEP=nccl $RUN 29544 bench.py $C > $O/bench_n8_cfg4_nccl.json 2> $O/bench_n8_cfg4_nccl.err
TTLDA_MSTEP=replicated $RUN 29545 bench.py $C > $O/bench_n8_cfg4_repl.json 2> $O/bench_n8_cfg4_repl.err
tail -c 300 $O/bench_n8_cfg4_*.err
