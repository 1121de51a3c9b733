# round 2, call H: mirror by bitmask counting sort — tests + e2e A/B against the CUB sort
set -x
O=gpurun_out/r02h
mkdir -p $O
python -m pytest tests -m gpu -q -k "csc or corpus or host or block or bow or cfg2 or cfg4 or cfg5 or readme" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sel.log
tail -6 $O/pytest_sel.log | cut -c1-200
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check"
$B > $O/bench_cfg3.json 2> $O/bench_cfg3.err
TTLDA_CSC_SORT=1 $B > $O/bench_cfg3_cubsort.json 2> $O/bench_cfg3_cubsort.err
python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu --no-extra --no-fit-check > $O/bench_cfg4.json 2> $O/bench_cfg4.err
TTLDA_CSC_SORT=1 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu --no-extra --no-fit-check > $O/bench_cfg4_cubsort.json 2> $O/bench_cfg4_cubsort.err
tail -c 300 $O/*.err
