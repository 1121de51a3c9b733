# round 2, call B: full GPU test suite + the default bench line (with extra configs) + the reference arm
set -x
O=gpurun_out/r02b
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench rc=$?"
tail -c 600 $O/bench_cfg3.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
ls -la $O
