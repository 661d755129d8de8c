run() { echo "== $*"; env "$@" timeout 200 python tools/headline_time.py 2>&1 | grep "refine" | tail -1; }
run A=1
run BBO_TF32_PANEL_GROUP_MB=8
run BBO_TF32_PANEL_GROUP_MB=16
run BBO_TF32_PANEL_GROUP_MB=2
run BBO_TF32_KS_BUDGET_MB=1200
run BBO_TF32_KS_BUDGET_MB=1536
run BBO_TF32_KS_BUDGET_MB=3072
sw() { echo "== sweep $*"; env "$@" timeout 200 python tools/n_sweep.py 6144 8192 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['n'], '%.4g %.4g'%(r['tf32_refine_cand_per_s'], r['tf32_cand_per_s']))"; }
sw A=1
sw BBO_TF32_PERSIST_MAX_MB=40
sw BBO_TF32_PERSIST_MAX_MB=200
