cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-ncu}; mkdir -p $O
python tools/prof_step.py bf16 > $O/plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"reduce_planes|morph_fused|tile_quantize" --launch-skip 54 --launch-count 18 -f -o /tmp/step_bf16 python tools/prof_step.py bf16 > $O/ncu_step.log 2>&1
ncu -i /tmp/step_bf16.ncu-rep --page raw --csv > $O/step_bf16_raw.csv 2>/dev/null
ls -la $O /tmp/step_bf16.ncu-rep
