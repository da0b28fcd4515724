cd /root/repo
for lib in - tools/_variants/lib_mb5.so tools/_variants/lib_mb6.so; do
timeout 300 python tools/tune_pcg.py $lib 16384 40 2>&1 | tail -1
timeout 300 python tools/tune_pcg.py $lib 4096 60 2>&1 | tail -1
done
