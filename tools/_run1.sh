python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_cnvt.py -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-clustered"
for u in 4 2 1; do $B --opt tile_fill_unroll=$u > gpurun_out/r2s2_fill$u.json 2>/dev/null; done
$B --opt tile_tma=0 > gpurun_out/r2s2_tma0.json 2>/dev/null
python - <<EOF
import json
for f in ["fill4","fill2","fill1","tma0"]:
    d=json.load(open("gpurun_out/r2s2_%s.json"%f)); print(f, d["ms_per_step"], d["stages_ms"]["sort"], d["stages_ms"]["assign"])
EOF
