# round 2, call C: GPU test suite + cfg5 shard with the topic-windowed statistics pass
set -x
O=gpurun_out/r02c
mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
python bench.py --config cfg5 --docs 1250000 --steps 3 --warmup 3 --no-cpu --no-fit-check > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench rc=$?"
TTLDA_KWINDOW=64 python bench.py --config cfg5 --docs 1250000 --steps 3 --warmup 3 --no-cpu --no-fit-check > $O/bench_cfg5_kw64.json 2> $O/bench_cfg5_kw64.err
TTLDA_KWINDOW=128 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu --no-fit-check > $O/bench_cfg4_kw128.json 2> $O/bench_cfg4_kw128.err
tail -c 400 $O/*.err
