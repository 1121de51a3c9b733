# round 2, call Q: slice-walk mirror v3.2 (16-byte cp.async fill) + the evidence of the final state: full GPU test suite, the
# default bench line, launch list of the default bench command, per-kernel split of the mirror
set -x
O=gpurun_out/r02q
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slice_walk" > $O/pytest_sel.log 2>&1
RC=$?; echo "pytest rc=$RC" >> $O/pytest_sel.log; tail -5 $O/pytest_sel.log | cut -c1-300
if [ $RC -ne 0 ]; then export TTLDA_CSC_WALK=0; echo "WALK PATH DISABLED FOR THE REST OF THE CALL"; fi
python tools/diag_csc.py > $O/diag_csc_walk.txt 2>&1
TTLDA_CSC_WALK=0 python tools/diag_csc.py > $O/diag_csc_cubsort.txt 2>&1
cat $O/diag_csc_walk.txt $O/diag_csc_cubsort.txt | tail -4
DIAG_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:csc_' -s 5 -c 12 --csv --log-file $O/launches_csc.csv python tools/diag_csc.py > $O/ncu_csc.log 2>&1
timeout 600 python tools/diag_e2e.py > $O/diag_e2e_walk.txt 2>&1; tail -7 $O/diag_e2e_walk.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log | cut -c1-300
timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
tail -c 300 $O/bench_default.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --no-fit-check"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_bench_cfg3.csv $CMD > $O/ncu_launches.log 2>&1
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02q/bench_default.json").read().strip().splitlines()[-1])
print("ms/step", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], j["e2e"]["int32"]["ms_per_step"], "frac", j["roofline"]["frac"])
PY
