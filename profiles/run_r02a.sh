# round 2, step A: FP64 analysis in the round trip
python -m pytest tests/test_gpu_roundtrip.py tests/test_gpu_stft.py tests/test_gpu_vocoder.py -x -q -m gpu > gpurun_out/pytest_a.log 2>&1; tail -15 gpurun_out/pytest_a.log
python profiles/roundtrip_error_vs_length.py > gpurun_out/rt_err_f64.txt 2>&1; cat gpurun_out/rt_err_f64.txt
CA_RT_F32_ANALYSIS=1 python profiles/roundtrip_error_vs_length.py > gpurun_out/rt_err_f32.txt 2>&1; cat gpurun_out/rt_err_f32.txt
CA_RT_TIMING=1 python profiles/time_roundtrip.py 32 60 2>&1 | tail -8
