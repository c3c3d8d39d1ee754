python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "owner or config2_at_512" 2>&1 | tail -2
python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for o in 1 0; do
  $B --opt tile_order=$o > gpurun_out/r2s2_v.json 2>gpurun_out/r2s2_v.err || tail -3 gpurun_out/r2s2_v.err
  python - <<EOF
import json
d=json.load(open("gpurun_out/r2s2_v.json")); print("order=$o", d["ms_per_step"], d["stages_ms"]["sort"], d["stages_ms"]["assign"], d["stages_ms"]["bin"], "clustered", d["clustered"]["ms_per_step"], d["clustered"]["stages_ms"]["sort"], d["clustered"]["stages_ms"]["assign"])
EOF
done
ncu --set full --clock-control none --import-source on -k regex:"k_tile_acc" -s 1 -c 1 -o gpurun_out/r2s2_prof4 $B --no-clustered --steps 1 > gpurun_out/p.log 2>&1
