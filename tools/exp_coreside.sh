# Co-residency sweep: K2 CTA size (MCAQ_K2_THREADS) x K3 register cap (K3_MINB: 3 -> 80 regs, 4 -> 64 regs), whole step
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-cores}; mkdir -p $O
for MB in 3 4; do
  touch mcaq_yolo_b200/csrc/tile_quantize.cu
  MCAQ_NVCC_EXTRA="-DK3_MINB=$MB" python mcaq_yolo_b200/build.py --force > /dev/null 2>&1
  for TH in 512 384 256; do
    for WL in yolov8n_640_b64_bf16; do
      MCAQ_K2_THREADS=$TH python bench.py --workload $WL --no-cpu-baseline --no-secondary --steps 100 --warmup 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('K3_MINB=$MB K2_THREADS=$TH $WL step', d['ms_per_step'], d['roofline']['whole_step']['frac'], 'serial', d['roofline']['serial_hook']['ms_per_forward'])"
    done
  done
done > $O/coreside.log 2>&1
python mcaq_yolo_b200/build.py --force > /dev/null 2>&1
cat $O/coreside.log
