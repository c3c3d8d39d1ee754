T="timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; np=$2; port=$3; shift 3
  $T --nproc-per-node $np --master-port $port bench.py --gpus $np "$@" > gpurun_out/r2s3_$name.json 2> gpurun_out/r2s3_$name.err
  python - <<EOF
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2s3_$name.json") if l.startswith("{")][-1])
    print("$name", round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stages_ms_max_over_ranks"].items()}, "nvlink", round(d["nvlink"]["frac"],3), d["nvlink"]["transport"][:20], "parity", d["parity"] and d["parity"]["ok"])
except Exception as ex:
    print("$name FAILED", ex); print(open("gpurun_out/r2s3_$name.err").read()[-800:])
EOF
}
run c2_4gpu_auto 4 29503 --steps 5 --warmup 3 --no-replicas --no-e2e
POWSPEC_B200_P2P=0 run c2_2gpu_nccl 2 29504 --steps 5 --warmup 3 --no-replicas --no-e2e
