cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-m}; mkdir -p $O
python bench.py --workload model_v8n_640_b64 --steps 10 > $O/model_v8n_640_b64.json 2> $O/model_v8n.err; tail -2 $O/model_v8n.err; cat $O/model_v8n_640_b64.json
python bench.py --workload model_v8s_1280_b32 --steps 5 > $O/model_v8s_1280_b32.json 2> $O/model_v8s.err; tail -2 $O/model_v8s.err; cat $O/model_v8s_1280_b32.json
python bench.py --workload model_v8n_640_b1 --steps 20 > $O/model_v8n_640_b1.json 2> $O/model_b1.err; cat $O/model_v8n_640_b1.json
