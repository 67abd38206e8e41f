cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-tr}; mkdir -p $O
python -m pytest tests/test_gpu_train_nets.py tests/test_gpu_train.py tests/test_gpu_modules.py -q -m gpu --tb=short 2>&1 | grep -E "^E  |what|Mismatch|Max |FAILED|passed|failed" | head -60
python tools/train_bench.py --dtype bf16 --out $O/train_b16_bf16.json > $O/train_bf16.log 2>&1; python -c "
import json; d=json.load(open('$O/train_b16_bf16.json')); print({k:v for k,v in d.items() if k!='kernels'})"
python tools/prof_train.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_launches.csv python tools/prof_train.py > $O/ncu_train.log 2>&1

python tools/sum_launches.py $O/train_launches.csv 4
