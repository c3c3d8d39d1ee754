(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-clustered"
for o in 1 0; do
  $B --opt fft_store_skip=$o > gpurun_out/r2s2_v.json 2>gpurun_out/r2s2_v.err || tail -3 gpurun_out/r2s2_v.err
  python - <<EOF
import json
d=json.load(open("gpurun_out/r2s2_v.json")); s=d["stages_ms"]; print("store_skip=$o", round(d["ms_per_step"],2), {k:round(v,2) for k,v in s.items() if v>0}, d["P0_first_bins"])
EOF
done
