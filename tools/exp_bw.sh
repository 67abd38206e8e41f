cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-bw}; mkdir -p $O
for SK in 2 3; do for KS in 3 4; do python tools/bw_stream_probe.py --skew $SK --k2-streams $KS 2>&1 | tail -1; done; done > $O/bw2.log 2>&1
cat $O/bw2.log
