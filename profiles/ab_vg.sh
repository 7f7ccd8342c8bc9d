# same-box A/B of the local-map VoxelGrid variants (bench.py C3, 20 steps each, twice)
run() { timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['stages_ms'].items() if v})"; }
for i in 1 2; do
run base
LVREG_LANE_PRIO=1 run prio_corner
LVREG_LANE_PRIO=2 run prio_surf
LVREG_LANE_PRIO=12 run prio_scans
LVREG_VG_BUCKET=0 run sortpath
done
LVREG_LANE_PRIO=1 LVREG_DEBUG_PHASES=1 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --sustain-seconds 0 2>&1 | grep -i "lane" | tail -3
