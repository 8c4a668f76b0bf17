timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for k in qkv fc2 qk sv; do timeout 100 python tools/profile_layer.py $k 1 2>&1 | tail -1; done
