# round 2, call D: GPU test suite + cfg3 bench with the hot-word cache (A/B against TTLDA_HOT=0) + window widths at cfg5
set -x
O=gpurun_out/r02d
mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check > $O/bench_cfg3_hot.json 2> $O/bench_cfg3_hot.err; echo "rc=$?"
TTLDA_HOT=0 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check > $O/bench_cfg3_nohot.json 2> $O/bench_cfg3_nohot.err
TTLDA_HOT_ROWS=128 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check > $O/bench_cfg3_hot128.json 2> $O/bench_cfg3_hot128.err
TTLDA_KWINDOW=256 python bench.py --config cfg5 --docs 1250000 --steps 3 --warmup 3 --no-cpu --no-fit-check > $O/bench_cfg5_kw256.json 2> $O/bench_cfg5_kw256.err
python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu --no-fit-check > $O/bench_cfg4.json 2> $O/bench_cfg4.err
python bench.py --config cfg2 --steps 20 --warmup 3 --no-cpu --no-fit-check > $O/bench_cfg2.json 2> $O/bench_cfg2.err
tail -c 300 $O/*.err
