T="timeout 400 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T --nproc-per-node 8 --master-port 29501 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2s2_c2_8gpu.json 2> gpurun_out/r2s2_c2_8gpu.err
$T --nproc-per-node 4 --master-port 29502 bench.py --gpus 4 --steps 5 --warmup 3 --no-replicas > gpurun_out/r2s2_c2_4gpu.json 2> gpurun_out/r2s2_c2_4gpu.err
$T --nproc-per-node 8 --master-port 29503 bench.py --gpus 8 --workload c4 --steps 2 --warmup 3 > gpurun_out/r2s2_c4_8gpu.json 2> gpurun_out/r2s2_c4_8gpu.err
$T --nproc-per-node 8 --master-port 29504 bench.py --gpus 8 --workload c5 --precision 4 --steps 2 --warmup 3 > gpurun_out/r2s2_c5_8gpu.json 2> gpurun_out/r2s2_c5_8gpu.err
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -m gpu -q 2>&1 | tail -4
for f in c2_8gpu c2_4gpu c4_8gpu c5_8gpu; do echo $f; tail -c 600 gpurun_out/r2s2_$f.json; tail -3 gpurun_out/r2s2_$f.err; done
