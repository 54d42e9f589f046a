for cfg in "4 4 2e-10" "2 4 2e-10" "1 4 2e-10" "4 2 2e-10" "2 2 2e-10" "4 4 1e-9" "2 2 1e-9" "1 1 1e-9" "1 1 3e-9"; do
  set -- $cfg
  echo "kappa=$1 qsafe=$2 tol=$3"
  QDSIM_E2_KAPPA=$1 QDSIM_E2_QSAFE=$2 QDSIM_E2_TOL=$3 timeout 300 python tools/tunnel_ab.py --cases 4:128,8:32 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  N=%d noda %.2f ms (hh %.2f) max_abs %.2e  >1e-9 %.2e  >1e-7 %.2e n>1e-6 %d'%(d['n_dot'],d['ms_noda'],d['ms_householder'],d['max_abs'],d['frac_gt_1e-9'],d['frac_gt_1e-7'],d['n_gt_1e-6']))
"
done
