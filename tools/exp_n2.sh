cd $GRAFT_REPO_ROOT
N=${2:-2}
O=gpurun_out/${1:-n2}; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 50 --warmup 10 > $O/bench_n$N.json 2> $O/bench_n$N.err; tail -3 $O/bench_n$N.err
python -c "
import json; d=json.load(open('$O/bench_n$N.json')); print({k: d.get(k) for k in ('value','ms_per_step','parity_check','n_gpus')}, d['e2e']['value'], d['e2e'].get('per_gpu'), d['e2e'].get('host'))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 tools/train_dist.py --out $O/train_n$N.json 2> $O/train_n$N.err; tail -3 $O/train_n$N.err
python tools/train_dist.py --out $O/train_n1.json 2> $O/train_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/pcie_probe.py > $O/pcie_n$N.json 2> $O/pcie_n$N.err; cat $O/pcie_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus $N --steps 50 --warmup 10 --workload yolov8s_1280_b32_f32 > $O/bench_v8s_n$N.json 2> $O/bench_v8s_n$N.err
python -c "
import json; d=json.load(open('$O/bench_v8s_n$N.json')); print('v8s', {k: d.get(k) for k in ('value','ms_per_step','parity_check','n_gpus')}, d['roofline']['whole_step']['frac'])"
