for ny in 32 64 128 256 512; do for cl in 8 1; do echo "== ny $ny cl $cl"; timeout 60 tools/wavebench 4096 $ny 2 50 16 $cl 2>&1 | tail -2; done; done > gpurun_out/wb3.log 2>&1
for ny in 32 64 256; do echo "== sigma3 ny $ny cl 8"; timeout 60 tools/wavebench 4096 $ny 3 50 16 8 2>&1 | tail -2; done >> gpurun_out/wb3.log 2>&1
for ny in 32 64 256; do echo "== subs8 ny $ny cl 8"; timeout 60 tools/wavebench 4096 $ny 2 50 8 8 2>&1 | tail -2; done >> gpurun_out/wb3.log 2>&1
cat gpurun_out/wb3.log
