# usage: quick_check.sh [layers...]   -- GPU tests, then per-launch sweep timing of the given layers (direct W search)
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
export P4V_PROFILE_LOG=1
for k in "${@:-qkv}"; do
  P4V_GRAM=0 timeout 120 python tools/profile_layer.py $k 1 2>&1 | grep -E "p4v sweep|rounds=" | awk '/rounds=/{print; next} {k=$5" "$6" "$7" "$10; n[k]++; s[k]+=$3; c[k]+=substr($NF, index($NF,"=")+1)} END{for(k in n) printf "%4d x %-60s avg %8.1f us  %6.0f cyc/acc\n", n[k], k, s[k]/n[k], c[k]/n[k]}'
done
