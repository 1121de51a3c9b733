# round 2, call E (2 GPUs): N>1 parity under pytest, the N=2 bench line (weak cfg3 + strong cfg4) in the three
# statistics-exchange modes, hot-cache / sharded-M-step kernel tests
set -x
O=gpurun_out/r02e
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
python -m pytest tests/test_named_configs.py tests/test_gpu_parity.py -m gpu -q -k "two_gpus or hot_word or non_current or malformed or streamed or partial_fit_keeps or cfg2_full or sharded_mstep" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sel.log
tail -8 $O/pytest_sel.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$RUN 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2_peer.json 2> $O/bench_n2_peer.err; echo "bench rc=$?"
tail -c 1500 $O/bench_n2_peer.err
TTLDA_MSTEP=nccl $RUN 29534 bench.py --gpus 2 --steps 5 --warmup 3 --no-parity > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err
TTLDA_MSTEP=replicated $RUN 29535 bench.py --gpus 2 --steps 5 --warmup 3 --no-parity > $O/bench_n2_repl.json 2> $O/bench_n2_repl.err
tail -c 600 $O/bench_n2_nccl.err $O/bench_n2_repl.err
