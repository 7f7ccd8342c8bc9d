# same-box A/B of the local-map VoxelGrid variants (bench.py C3, 20 timed steps each, twice):
#   base        sample-sort path (voxelgrid_bucket.cuh), packed coordinates for the sample / split kernels
#   nowkey      the same with keys computed from the points everywhere
#   cachedfirst the map lanes' chains enqueued before the scan lanes
#   sortpath    device-wide radix sort over the same voxel-ordered cache
run() { timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['stages_ms'].items() if v})"; }
for i in 1 2; do
run base
LVREG_DEBUG_NOWKEY=1 run nowkey
LVREG_VG_CACHED_FIRST=1 run cachedfirst
LVREG_VG_BUCKET=0 run sortpath
done
