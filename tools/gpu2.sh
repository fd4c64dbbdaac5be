for a in "4096 4096 2 5 16 8 1" "1000 777 2 5 16 2 1"; do echo "== $a"; timeout 60 tools/wavebench $a 2>&1 | tail -2; done
