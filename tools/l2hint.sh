for h in 0 1 2 3; do
  KOVER_B200_L2HINT=$h VARIANTS="1" bash tools/variants.sh 2>&1 | grep -v Warn | sed "s/^/l2hint=$h /"
done
