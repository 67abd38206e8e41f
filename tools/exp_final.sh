cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-final2}; mkdir -p $O
python -m pytest tests -q -m gpu > $O/gputest.log 2>&1; tail -3 $O/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > $O/bench_b64_bf16.json 2> $O/bench.err; tail -2 $O/bench.err
python -c "
import json; d=json.load(open('$O/bench_b64_bf16.json')); r=d['roofline']
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', r['frac'], 'serial', r['serial_hook'], 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline'), 'secondary', d.get('secondary', {}).get('roofline_frac_whole_step'), 'clocks', d['clocks'])"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_ref.err; cut -c1-600 $O/bench_reference.json
