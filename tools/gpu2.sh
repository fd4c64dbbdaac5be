for a in "4096 32 2 3 16 8" "4096 32 2 3 8 8" "4096 64 2 3 16 8" "4096 4096 2 3 16 8" "4096 4096 2 3 8 8"; do echo "== $a"; timeout 60 tools/wavebench_p $a 2>&1 | grep -v "dirty\|bwd cells"; done > gpurun_out/wb6.log 2>&1
cat gpurun_out/wb6.log
