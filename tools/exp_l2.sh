# L2 residency hints: K1 loads x evict_last, K3 reads it evict_first; whole step at several in-flight depths
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-l2}; mkdir -p $O
for V in "-DMCAQ_L2_HINTS=0" "-DMCAQ_L2_HINTS=1"; do
  echo "=== $V"
  MCAQ_NVCC_EXTRA="$V" python mcaq_yolo_b200/build.py --force > /dev/null 2>&1
  for F in 1 2 4; do
    python bench.py --no-cpu-baseline --no-secondary --steps 100 --warmup 10 --inflight $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('inflight $F step', d['ms_per_step'], d['roofline']['whole_step']['frac'], 'serial', d['roofline']['serial_hook']['ms_per_forward'])"
  done
done > $O/l2_hints.log 2>&1
python mcaq_yolo_b200/build.py --force > /dev/null 2>&1
cat $O/l2_hints.log
