for a in "4096 4096 2 5 16 2 1" "4096 4096 2 5 16 4 1" "4096 4096 2 20 16 8 1" "4096 4096 2 20 16 8 0" "1000 777 2 5 16 2 1" "2048 4096 3 5 16 4 1" "4096 2048 2 5 8 2 1" "777 3000 2 5 16 8 1"; do echo "== $a"; timeout 60 tools/wavebench $a 2>&1 | tail -4; done > gpurun_out/wb8.log 2>&1
cat gpurun_out/wb8.log | grep -v "dirty polled slots after the run: 0\|bwd cells with |err| > 1e-9: 0"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
