for v in "B2_TREE_LANES=2" "B2_TREE_LANES=4" "B2_TREE_LANES=8" "B2_NO_TREE_LANES=1"; do
echo "== $v"
env $v EXP_CAPS=100 python tools/exp_pgs.py c3 2>&1 | tail -1
env $v EXP_CAPS=100 python tools/exp_pgs.py c5 2>&1 | tail -1
done
