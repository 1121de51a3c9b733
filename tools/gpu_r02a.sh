# round 2, call A: bench lines + ncu --set full captures of the two hot kernels at the HBM-class shapes (cfg4, cfg5 shard)
set -x
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv
for CFG in "cfg4" "cfg5 --docs 1250000"; do
  T=$(echo $CFG | cut -d' ' -f1)
  python bench.py --config $CFG --steps 3 --warmup 3 --no-cpu > $O/bench_$T.json 2> $O/bench_$T.err &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:doc_pass_kernel|sstats_csc_kernel' -s 5 -c 2 -f -o /tmp/prof_$T python bench.py --config $CFG --steps 3 --warmup 3 --no-cpu > $O/ncu_$T.log 2>&1
  ncu -i /tmp/prof_$T.ncu-rep --page raw --csv > $O/raw_$T.csv 2>/dev/null
done
ls -la $O
