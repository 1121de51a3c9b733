# round 2, call G: GPU test suite; cfg3 bench: default / hot-word cache (LDS+LDG form) / G=16 lanes per row / CUB-sort mirror
set -x
O=gpurun_out/r02g
mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log | cut -c1-200
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check"
$B > $O/bench_cfg3.json 2> $O/bench_cfg3.err
TTLDA_CSC_SORT=1 $B > $O/bench_cfg3_cubsort.json 2> $O/bench_cfg3_cubsort.err
TTLDA_HOT=1 $B > $O/bench_cfg3_hot.json 2> $O/bench_cfg3_hot.err
TTLDA_HOT=1 TTLDA_HOT_ROWS=128 $B > $O/bench_cfg3_hot128.json 2> $O/bench_cfg3_hot128.err
TTLDA_GROUP=16 $B > $O/bench_cfg3_g16.json 2> $O/bench_cfg3_g16.err
tail -c 300 $O/*.err
