# K2 iteration loop on the GPU box: parity tests, K2 latency per shape, warp-instruction counts, bench line
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-k2}; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/gputest.log 2>&1; tail -4 $O/gputest.log
python tools/k2_bench.py 64 1 2>&1 | grep -v "C=256 H= 80\|C=512" > $O/k2_bench.log; cat $O/k2_bench.log
for shp in "64 80" "128 40" "256 20"; do
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:morph_fused --launch-skip 2 --launch-count 1 python tools/prof_k2.py $shp 64 2>&1 | grep -E "inst_executed|time_duration" ; done > $O/k2_inst.log 2>&1; cat $O/k2_inst.log
python bench.py --no-cpu-baseline --steps 100 --warmup 10 2>/dev/null > $O/bench.json; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['whole_step']['frac'], d['roofline']['kernel_ms'])"
if [ -n "$2" ]; then ncu --set full --clock-control none --import-source on -k regex:morph_fused --launch-skip 2 --launch-count 1 -f -o $O/k2_c3 python tools/prof_k2.py 64 80 64 > $O/ncu_c3.log 2>&1; fi
