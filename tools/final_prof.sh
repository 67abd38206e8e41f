set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/final
mkdir -p $O
python bench.py > $O/bench_b64_bf16.json 2> $O/bench.err
python bench.py --workload yolov8s_1280_b32_f32 > $O/bench_v8s1280_b32_f32.json 2> $O/bench_v8s.err
python tools/train_bench.py --dtype bf16 --out $O/train_b16_bf16.json > $O/train_bf16.log 2>&1
python tools/train_bench.py --dtype f32 --out $O/train_b16_f32.json > $O/train_f32.log 2>&1
python tools/op_sweep.py --out $O/op_sweep.jsonl > $O/op_sweep.log 2>&1
for d in bf16 f32; do python tools/kernel_bench.py --dtype $d; done > $O/kernel_bench.log 2>&1
python tools/stage_clocks.py 64 > $O/stage_clocks.log 2>&1
python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --inflight 1 > $O/plain_eager.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_eager.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --inflight 1 > $O/ncu_launches.log 2>&1
python tools/prof_step.py bf16 > $O/plain_step.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"reduce_planes|morph_fused|tile_quantize" --launch-skip 18 --launch-count 9 -f -o $O/step_bf16 python tools/prof_step.py bf16 > $O/ncu_step.log 2>&1
python tools/prof_step.py f32 > $O/plain_step_f32.log 2>&1 && \
ncu --set full --clock-control none -k regex:"reduce_planes|tile_quantize" --launch-skip 12 --launch-count 6 -f -o $O/k1k3_f32 python tools/prof_step.py f32 > $O/ncu_f32.log 2>&1
ls -la $O
