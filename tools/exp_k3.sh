# K3 register / unroll variants: rebuild tile_quantize.cu on the box per variant, time K1 / K3 in isolation
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-k3}; mkdir -p $O
for V in "-DK3_UNROLL=8 -DK3_MINB=3" "-DK3_UNROLL=4 -DK3_MINB=3" "-DK3_UNROLL=4 -DK3_MINB=4" "-DK3_UNROLL=8 -DK3_MINB=2"; do
  echo "=== $V"
  touch mcaq_yolo_b200/csrc/tile_quantize.cu
  MCAQ_NVCC_EXTRA="$V" python mcaq_yolo_b200/build.py --force > /dev/null 2>&1
  python tools/kernel_bench.py --dtype bf16 2>&1 | grep "K3 tile_quantize"
  python tools/kernel_bench.py --dtype f32 2>&1 | grep "K3 tile_quantize"
  python bench.py --no-cpu-baseline --steps 100 --warmup 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step', d['ms_per_step'], d['roofline']['whole_step']['frac'], 'serial', d['roofline']['serial_hook']['ms_per_forward'])"
done > $O/k3_variants.log 2>&1
cat $O/k3_variants.log
