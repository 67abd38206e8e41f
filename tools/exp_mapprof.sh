cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-mp}; mkdir -p $O
ncu --set full --import-source on --clock-control none -k regex:mapper_train --launch-skip 18 --launch-count 6 -f -o $O/map python tools/prof_train.py > $O/ncu.log 2>&1
ncu -i $O/map.ncu-rep --page source --csv --print-source cuda,sass > $O/map_source.csv 2>/dev/null
ncu -i $O/map.ncu-rep --page raw --csv > $O/map_raw.csv 2>/dev/null
rm -f $O/map.ncu-rep
ls -la $O
