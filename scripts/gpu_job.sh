# round-2 GPU job 66: repeat-and-compare stress of the tower's hand-over
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nnet_gpu.py -x -q --timeout=300 --timeout-method=thread -k "hand_over" 2>&1 | tail -3
