# K3 channels-per-CTA sweep (MCAQ_K3_CHUNK = 8 / 16 / 32): isolated K3 and the whole step
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-k3c}; mkdir -p $O
for CH in 16 8; do
  echo "=== MCAQ_K3_CHUNK=$CH"
  export MCAQ_K3_CHUNK=$CH
  python tools/kernel_bench.py --dtype bf16 2>&1 | grep "K3 tile_quantize"
  python tools/kernel_bench.py --dtype f32 2>&1 | grep "K3 tile_quantize"
  for WL in yolov8n_640_b64_bf16 yolov8s_1280_b32_f32; do
  python bench.py --workload $WL --no-cpu-baseline --no-secondary --steps 100 --warmup 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$WL step', d['ms_per_step'], d['roofline']['whole_step']['frac'], 'serial', d['roofline']['serial_hook']['ms_per_forward'])"
  done
done > $O/k3_chunk.log 2>&1
cat $O/k3_chunk.log
