python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "owner" 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-clustered"
$B > gpurun_out/r2s2_b.json 2>/dev/null
python - <<EOF
import json
d=json.load(open("gpurun_out/r2s2_b.json")); print(d["ms_per_step"], d["stages_ms"])
EOF
