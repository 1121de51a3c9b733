# round 2, call R: ncu --set full + source pages of the mirror kernels (v3.2)
set -x
O=gpurun_out/r02r
mkdir -p $O
DIAG_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:csc_walk_kernel|csc_walk_place' -s 3 -c 3 -f -o /tmp/prof_csc python tools/diag_csc.py > $O/ncu_csc.log 2>&1
ncu -i /tmp/prof_csc.ncu-rep --page raw --csv > $O/raw_csc.csv 2>/dev/null
ncu -i /tmp/prof_csc.ncu-rep --page source --csv -k regex:csc_walk_kernel > $O/source_csc_walk.csv 2>/dev/null
ncu -i /tmp/prof_csc.ncu-rep --page source --csv -k regex:csc_walk_place > $O/source_csc_place.csv 2>/dev/null
ls -la $O
