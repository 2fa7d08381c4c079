python -m pytest tests/test_gpu_classic.py -x -q -k "registered or zero_copy or invalid" 2>&1 | tail -3
python tools/pcie_probe.py 2>&1 | tail -7
