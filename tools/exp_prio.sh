cd $GRAFT_REPO_ROOT
run() { python bench.py --no-cpu-baseline --steps 100 --warmup 10 "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['whole_step']['frac'], 'serial', d['roofline']['serial_hook']['ms_per_forward'], d['roofline']['kernel_ms'])"; }
echo "default"; run
echo "v8s 1280 f32"; run --workload yolov8s_1280_b32_f32
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -2
