python bench.py > gpurun_out/r2s2_final1.json 2> gpurun_out/r2s2_final1.err; tail -2 gpurun_out/r2s2_final1.err
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-clustered"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s2_launches_final.csv $B > gpurun_out/l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fft_strided|k_spectrum|k_tile" -s 14 -c 8 -o gpurun_out/r2s2_prof5 $B > gpurun_out/p.log 2>&1
python - <<EOF
import json
d=json.load(open("gpurun_out/r2s2_final1.json")); print(d["ms_per_step"], d["stages_ms"], d["e2e"]["ms_per_step"], d["e2e"]["e2e_cold_ms"], d["parity"], d["clustered"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["stage"])
EOF
