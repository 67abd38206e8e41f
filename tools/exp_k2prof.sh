# ncu --set full of ONE steady-state K2 launch (C3 of step 4) with source correlation; CSV exports made on the box
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-k2p}; mkdir -p $O
python tools/prof_step.py bf16 > $O/plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:morph_fused --launch-skip ${2:-12} --launch-count 1 -f -o $O/k2 python tools/prof_step.py bf16 > $O/ncu.log 2>&1
ncu -i $O/k2.ncu-rep --page source --csv --print-source cuda,sass > $O/k2_source.csv 2>/dev/null
ncu -i $O/k2.ncu-rep --page raw --csv > $O/k2_raw.csv 2>/dev/null
ls -la $O
