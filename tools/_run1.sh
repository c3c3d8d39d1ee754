python -m pytest tests/test_dropin_binary.py -m gpu -x -q 2>&1 | tail -5
PSB_TRACE=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-clustered 2>&1 >/dev/null | grep -v "device " | head -24
