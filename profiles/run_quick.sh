# quick GPU check used while tuning
python -m pytest tests/test_gpu_oscbank.py tests/test_gpu_roundtrip.py -x -q -m gpu > gpurun_out/pytest_q.log 2>&1; tail -3 gpurun_out/pytest_q.log
python profiles/time_osc.py 2>&1 | tail -4
CA_RT_TIMING=1 python profiles/time_roundtrip.py 32 60 2>&1 | tail -12
