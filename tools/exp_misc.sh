cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-misc}; mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/gputest.log 2>&1; tail -3 $O/gputest.log
python tools/k2_bench.py 64 1 2>&1 | grep -v "C=256 H= 80\|C=512\|H=160"
python tools/nhwc_bench.py 2>&1 | tail -12
python tools/score_bench.py --ref > $O/score_bench.jsonl 2> $O/score.err; cat $O/score_bench.jsonl; tail -2 $O/score.err
python bench.py --no-cpu-baseline --no-secondary --steps 100 --warmup 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['whole_step']['frac'], d['roofline']['serial_hook']['ms_per_forward'], d['roofline']['kernel_ms'])"
