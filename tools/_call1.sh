bash tools/gpu.sh check e1
bash tools/gpu.sh numpy e1
bash tools/gpu.sh ncu e1
bash tools/gpu.sh ncu-tc e1
bash tools/gpu.sh bench e1
