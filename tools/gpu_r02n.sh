# round 2, call N: where does the slice-walk mirror spend its time?  per-kernel ncu --set full + source page of the walk
set -x
O=gpurun_out/r02n
mkdir -p $O
python tools/diag_csc.py > $O/diag_csc_walk.txt 2>&1
TTLDA_CSC_WALK=0 python tools/diag_csc.py > $O/diag_csc_cubsort.txt 2>&1
cat $O/diag_csc_walk.txt $O/diag_csc_cubsort.txt | tail -4
DIAG_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:csc_' -s 5 -c 10 -f -o /tmp/prof_csc python tools/diag_csc.py > $O/ncu_csc.log 2>&1
ncu -i /tmp/prof_csc.ncu-rep --page raw --csv > $O/raw_csc.csv 2>/dev/null
ncu -i /tmp/prof_csc.ncu-rep --page source --csv -k regex:csc_walk_kernel > $O/source_csc_walk.csv 2>/dev/null
ncu -i /tmp/prof_csc.ncu-rep --page source --csv -k regex:csc_walk_place > $O/source_csc_place.csv 2>/dev/null
ls -la $O
