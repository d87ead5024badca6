# round-2 GPU job 64: tower equality at more sizes (unit-geometry edges)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nnet_gpu.py -x -q --timeout=300 --timeout-method=thread -k "tower_implementations or slot_count" 2>&1 | tail -3
