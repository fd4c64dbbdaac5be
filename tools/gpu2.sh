for a in "4096 32 1 5 16 8" "4096 32 2 5 16 8" "4096 4096 1 5 16 8" "4096 4096 1 5 8 8" "4096 4096 2 5 16 8"; do echo "== $a"; timeout 60 tools/wavebench_p $a 2>&1 | grep -v "dirty\|bwd cells"; done > gpurun_out/wb7.log 2>&1
for a in "1000 777 1 5 16 8" "130 40 1 5 8 8" "16384 2048 1 5 16 8" "4096 4096 1 20 16 1" "4096 4096 3 20 16 8"; do echo "== $a"; timeout 60 tools/wavebench $a 2>&1 | tail -2; done >> gpurun_out/wb7.log 2>&1
cat gpurun_out/wb7.log
