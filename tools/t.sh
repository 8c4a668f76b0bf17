timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
P4V_TORCH_PROF=1 timeout 100 python tools/profile_layer.py qkv 1 2>&1 | grep -E "rounds=|gram_update|gram_eval|total device"
