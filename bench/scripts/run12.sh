for i in 1 2; do
BMPC_MSM_SUBWINDOWS=1 python bench/prove_ab.py 22 6 2>&1 | tail -1 | sed 's/^/subw=1 /'
BMPC_MSM_SUBWINDOWS=0 python bench/prove_ab.py 22 6 2>&1 | tail -1 | sed 's/^/subw=0 /'
done
BMPC_ACC_AFFINE=0 python bench/prove_ab.py 22 6 2>&1 | tail -1 | sed 's/^/xyzz   /'
