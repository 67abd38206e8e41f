cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-tma}; mkdir -p $O
for V in "-DMCAQ_TQ_CH=32 -DMCAQ_TQ_STAGES=2 -DMCAQ_TQ_MINB=3" "-DMCAQ_TQ_CH=16 -DMCAQ_TQ_STAGES=3 -DMCAQ_TQ_MINB=4" "-DMCAQ_TQ_CH=16 -DMCAQ_TQ_STAGES=4 -DMCAQ_TQ_MINB=3" "-DMCAQ_TQ_CH=32 -DMCAQ_TQ_STAGES=3 -DMCAQ_TQ_MINB=2"; do
  echo "=== $V"
  touch mcaq_yolo_b200/csrc/tile_quantize_tma.cu
  MCAQ_NVCC_EXTRA="$V" python mcaq_yolo_b200/build.py > /dev/null 2>&1
  timeout 200 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "tma and shape0" 2>&1 | tail -1
  for d in bf16 f32; do timeout 300 python tools/kernel_bench.py --dtype $d --only k3 2>&1 | grep -E "K3 ranges"; done
done > $O/k3_tma_variants.log 2>&1; cat $O/k3_tma_variants.log
