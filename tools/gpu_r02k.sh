# round 2, call K: the evidence captures of the final kernels — launch list of the default bench command, ncu --set full
# of one EM iteration at cfg3 / cfg4 / cfg5 shard (raw CSV pages come back; the .ncu-rep files stay on the box)
set -x
O=gpurun_out/r02k
mkdir -p $O
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --no-fit-check"
$CMD > $O/bench_cfg3_short.json 2> $O/bench_cfg3_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_cfg3.csv $CMD > $O/ncu_launches.log 2>&1
for CFG in "cfg3" "cfg4" "cfg5 --docs 1250000"; do
  T=$(echo $CFG | cut -d' ' -f1)
  C2="python bench.py --config $CFG --steps 2 --warmup 3 --no-cpu --no-extra --no-fit-check --no-e2e"
  $C2 > $O/bench_$T.json 2> $O/bench_$T.err &&
  timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:doc_pass|sstats_|theta_refresh_kernel|mstep_lambda|beta_refresh' -s 4 -c 64 -f -o /tmp/prof_$T $C2 > $O/ncu_$T.log 2>&1
  ncu -i /tmp/prof_$T.ncu-rep --page raw --csv > $O/raw_$T.csv 2>/dev/null
done
ncu -i /tmp/prof_cfg3.ncu-rep --page source --csv -k regex:doc_pass_kernel > $O/source_estep_cfg3.csv 2>/dev/null
ls -la $O; du -sh $O
