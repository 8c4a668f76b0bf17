timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for k in qkv fc2 sv; do timeout 100 python tools/profile_layer.py $k 3 2>&1 | tail -1; done
