timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grad or stored or config5 or slab" 2>&1 | tail -1
timeout 900 python tools/grad_bench.py 2>&1 | grep "stored"
