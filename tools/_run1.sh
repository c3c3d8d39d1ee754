(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3)
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-clustered"
for o in 1 0; do
  $B --opt tile_tma=$o > gpurun_out/r2s2_v.json 2>gpurun_out/r2s2_v.err || tail -3 gpurun_out/r2s2_v.err
  python - <<EOF
import json
d=json.load(open("gpurun_out/r2s2_v.json")); s=d["stages_ms"]; print("tma=$o", round(d["ms_per_step"],2), {k:round(v,2) for k,v in s.items() if v>0}, "e2e", round(d["e2e"]["ms_per_step"],2), "cold", round(d["e2e"]["e2e_cold_ms"],1))
EOF
done
