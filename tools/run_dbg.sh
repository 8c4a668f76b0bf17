export P4V_PROFILE_LOG=1 P4V_GRAM=0
for dbg in 0 1 2 3; do
  for op in bf16 int8; do
    echo "=== debug=$dbg operand=$op"
    P4V_SWEEP_DEBUG=$dbg P4V_OPERAND=$op timeout 120 python tools/profile_layer.py ${1:-qkv} 1 2>&1 | grep -E "p4v sweep" | awk '{k=$5" "$6" "$7" "$10; n[k]++; s[k]+=$3; c[k]+=substr($NF, index($NF,"=")+1)} END{for(k in n) printf "%4d x %-60s avg %8.1f us  %6.0f cyc/acc\n", n[k], k, s[k]/n[k], c[k]/n[k]}'
  done
done
