cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-ba}; mkdir -p $O
python -m pytest tests/test_backend_agreement.py -q -m gpu 2>&1 | tail -2
python tools/backend_agreement.py --synthetic 16 --imgsz 640 --json $O/backend_agreement.json 2>&1 | tail -12
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"reduce_planes|morph_fused|tile_quantize" -c 2000 --csv --log-file $O/launches_bench_kernels.csv $BENCH > $O/ncu_launches.log 2>&1
python tools/sum_launches.py $O/launches_bench_kernels.csv 1 | head -14
