P4V_MMA_WARPS=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "mbarrier wait" | tail -4
P4V_MMA_WARPS=2 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "mbarrier wait" | grep -E "^FAILED|passed|failed" | tail -4
