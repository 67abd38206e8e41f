# Round-2 evidence batch (one B200): GPU tests, smoke, both bench arms, ncu launch list of the bench command, ncu --set full of
# one steady-state step (CSV exports made on the box), isolated kernels, scoring, training step.  Outputs: gpurun_out/<tag>/
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-ev2}; mkdir -p $O
python -m pytest tests -q -m gpu > $O/gputest.log 2>&1; tail -2 $O/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench_b64_bf16.json 2> $O/bench.err; tail -2 $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_ref.err
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary"
$BENCH > $O/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"reduce_planes|morph_fused|tile_quantize" -c 2000 --csv --log-file $O/launches_bench.csv $BENCH > $O/ncu_launches.log 2>&1
python tools/prof_step.py bf16 > $O/plain_step.log 2>&1 && \
ncu --set full --clock-control none -k regex:"reduce_planes|morph_fused|tile_quantize" --launch-skip 36 --launch-count 9 -f -o /tmp/step_bf16 python tools/prof_step.py bf16 > $O/ncu_step.log 2>&1
ncu -i /tmp/step_bf16.ncu-rep --page raw --csv > $O/ncu_full_step_bf16_raw.csv 2>/dev/null
for d in bf16 f32; do python tools/kernel_bench.py --dtype $d; done > $O/kernel_bench.txt 2>&1
python tools/score_bench.py --ref > $O/score_bench.jsonl 2> $O/score.err
python tools/train_bench.py --dtype bf16 --out $O/train_b16_bf16.json > $O/train_bf16.log 2>&1
python tools/prof_train.py > $O/plain_train.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train.csv python tools/prof_train.py > $O/ncu_train.log 2>&1
ls -la $O
python -c "
import json; d=json.load(open('$O/bench_b64_bf16.json')); r=d['roofline']
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', r['frac'], 'serial', r['serial_hook'], 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline'), 'secondary', d.get('secondary', {}).get('roofline_frac_whole_step'), 'clocks', d['clocks'])"
cut -c1-500 $O/bench_reference.json
