python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "config2_at_512" 2>&1 | grep -E "config 2|passed|failed|rel err|Error" | head
bash tools/sanitize.sh gpurun_out 2>&1 | tail -40
