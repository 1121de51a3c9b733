# round 2, call I (2 GPUs): sharded M-step with NVSwitch multicast — parity under pytest, N=2 bench with / without it
set -x
O=gpurun_out/r02i
mkdir -p $O
python -m pytest tests/test_named_configs.py tests/test_gpu_parity.py -m gpu -q -k "two_gpus or sharded_mstep or counting_sort" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sel.log
tail -5 $O/pytest_sel.log | cut -c1-200
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
A="--gpus 2 --steps 5 --warmup 3 --no-e2e --no-fit-check"
$RUN 29533 bench.py $A > $O/bench_n2_mc.json 2> $O/bench_n2_mc.err; echo "rc=$?"
TTLDA_MULTICAST=0 $RUN 29534 bench.py $A --no-parity > $O/bench_n2_nomc.json 2> $O/bench_n2_nomc.err
tail -c 700 $O/bench_n2_mc.err
