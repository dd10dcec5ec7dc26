for it in 1622 1703 1803 1811 1838 2073; do timeout 60 python tests/stress_gpu.py --first $it --seconds 0.01 2>&1 | grep "FAIL" | cut -c1-400; done
python -m pytest tests/test_full_size_gpu.py -x -q -m gpu -k db_batch 2>&1 | grep "^E" | head -12
