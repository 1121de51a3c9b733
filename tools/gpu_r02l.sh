# round 2, call L: the slice-walk mirror (tests, e2e A/B against the CUB pair sort), TMA-fed E-step A/B, the default bench
# line, and the evidence captures of the final kernels (launch list of the default bench command, ncu --set full of one EM
# iteration at cfg3 / cfg4 / cfg5 shard; raw CSV pages come back, the .ncu-rep files stay on the box)
set -x
O=gpurun_out/r02l
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slice_walk" > $O/pytest_walk.log 2>&1
WRC=$?; echo "pytest rc=$WRC" >> $O/pytest_walk.log; tail -15 $O/pytest_walk.log | cut -c1-300
if [ $WRC -ne 0 ]; then export TTLDA_CSC_WALK=0; echo "WALK PATH DISABLED FOR THE REST OF THE CALL"; fi
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log | cut -c1-300
S="--steps 10 --warmup 3 --no-cpu --no-extra --no-fit-check"
timeout 600 python bench.py $S > $O/bench_short_walk.json 2> $O/bench_short_walk.err
TTLDA_CSC_WALK=0 timeout 600 python bench.py $S > $O/bench_short_cubsort.json 2> $O/bench_short_cubsort.err
TTLDA_TMA=1 timeout 600 python bench.py $S --no-e2e > $O/bench_short_tma.json 2> $O/bench_short_tma.err
timeout 600 python tools/diag_e2e.py > $O/diag_e2e_walk.txt 2>&1
TTLDA_CSC_WALK=0 timeout 600 python tools/diag_e2e.py > $O/diag_e2e_cubsort.txt 2>&1
timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
tail -c 400 $O/bench_default.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --no-fit-check"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_bench_cfg3.csv $CMD > $O/ncu_launches.log 2>&1
for CFG in "cfg3" "cfg4" "cfg5 --docs 1250000"; do
  T=$(echo $CFG | cut -d' ' -f1)
  C2="python bench.py --config $CFG --steps 2 --warmup 3 --no-cpu --no-extra --no-fit-check --no-e2e"
  timeout 600 $C2 > $O/bench_$T.json 2> $O/bench_$T.err &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:doc_pass|sstats_|theta_refresh_kernel|mstep_lambda|beta_refresh' -s 4 -c 48 -f -o /tmp/prof_$T $C2 > $O/ncu_$T.log 2>&1
  ncu -i /tmp/prof_$T.ncu-rep --page raw --csv > $O/raw_$T.csv 2>/dev/null
done
ncu -i /tmp/prof_cfg3.ncu-rep --page source --csv -k regex:doc_pass_kernel > $O/source_estep_cfg3.csv 2>/dev/null
ls -la $O; du -sh $O
