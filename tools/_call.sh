bash tools/gpu.sh check c4
bash tools/gpu.sh ppo c4f 130
bash tools/gpu.sh ppo c4t 45 --update torch --amp
bash tools/gpu.sh ncu-train c4
