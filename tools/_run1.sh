(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3)
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-clustered"
$B > gpurun_out/r2s2_v.json 2>gpurun_out/r2s2_v.err || tail -3 gpurun_out/r2s2_v.err
python - <<EOF
import json
d=json.load(open("gpurun_out/r2s2_v.json")); s=d["stages_ms"]
print(round(d["ms_per_step"],2), {k:round(v,2) for k,v in s.items() if v>0})
EOF
