cd /root/repo
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for f in 0 4; do timeout 200 python tools/tune_pcg.py - 16384 40 $f 2>&1 | tail -1; done | tee gpurun_out/r2_tune10.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
