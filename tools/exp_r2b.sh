set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2b; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for T in 256 384 512; do MCAQ_K2_THREADS=$T python tools/k2_bench.py 64 1,2 2>&1 | grep -v "H=160\|C=256 H= 80\|C=512"; done > $O/k2_threads.log 2>&1
for T in 0 384 512; do for F in 2 4 8; do echo "threads=$T inflight=$F"; MCAQ_K2_THREADS=$T python bench.py --no-cpu-baseline --steps 100 --warmup 10 --inflight $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['whole_step']['frac'], d['roofline']['kernel_ms'])"; done; done > $O/bench_threads.log 2>&1
for S in 2; do for F in 2 4; do echo "split=$S inflight=$F"; MCAQ_K2_SPLIT=$S python bench.py --no-cpu-baseline --steps 100 --warmup 10 --inflight $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['whole_step']['frac'])"; done; done >> $O/bench_threads.log 2>&1
cat $O/k2_threads.log $O/bench_threads.log
