for st in 0 4096 8192 10240 12288; do
echo "stage $st"; B2_PGS_STAGE=$st EXP_CAPS=100 python tools/exp_pgs.py c4 2>&1 | tail -1 | cut -c1-200
done
