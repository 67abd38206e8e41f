cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-t}; mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/gputest.log 2>&1; tail -25 $O/gputest.log
