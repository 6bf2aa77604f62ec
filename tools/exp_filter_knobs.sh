for env in "" "DSIR_TC_DEBUG=1" "DSIR_TC_PRIME=0" "DSIR_TC_PRIME=4" "DSIR_TC_PRIME=16" "DSIR_TC_DEBUG=1 DSIR_TC_PRIME=0"; do
  echo "== $env"; env $env timeout 120 python tools/time_kernels.py 2>&1 | grep -i "match_argmin\[tc\]\|filter kernel\|rescued"
done
